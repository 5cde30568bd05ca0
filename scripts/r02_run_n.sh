#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bandwidth.py tests/test_gpu_wrapper.py -m gpu -x -q -p no:cacheprovider --timeout 300 -k "adamw or wrapper or recipe" > gpurun_out/pytest_n.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_n.log)"
timeout 600 python bench.py --layers 8 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cfg4 --busy-trace > gpurun_out/bench_busy8c.json 2> gpurun_out/bench_busy8c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_busy8c.json'))
b=d['busy']; print(d['ms_per_step'], d['optimizer_ms'], b['busy_frac'])
for r in b['top'][:6]: print(f"{r['ms']:9.3f} ms {r['launches']:5d}  {r['kernel'][:60]}")
PY
