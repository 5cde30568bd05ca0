#!/bin/bash
# 4-GPU sanity of the data-parallel bench (the driver's scaling run uses the same command line)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
echo "bench4 exit=$?"; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_4gpu.json') if l.startswith('{')][-1])
    print(d['value'], d['ms_per_step'], d['n_gpus'], d['allreduce']['exposed_ms_per_step'], d['allreduce']['n_pieces'], d['cfg4'] and d['cfg4']['value'], d['e2e']['value'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_4gpu.err').read()[-1500:])
PY
