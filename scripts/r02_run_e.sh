#!/bin/bash
# Round-2 run e (2 GPUs): data-parallel bench with the overlapped gradient all-reduce; NCCL's own description of the
# communicator (algorithm / protocol / channels) and the exposed part of the reduction measured with CUDA events.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH,TUNING
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 4 --warmup 3 --no-cfg4 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench2 exit=$?"; tail -c 1500 gpurun_out/bench_2gpu.json
grep -E "NCCL INFO (Channel|Connected|comm 0x|Using|NVLS|Trees|Ring|P2P|ncclCommInitRank)" gpurun_out/bench_2gpu.json gpurun_out/bench_2gpu.err | cut -c1-220 | sort | uniq -c | sort -rn | head -40 > gpurun_out/nccl_2gpu_summary.txt
unset NCCL_DEBUG NCCL_DEBUG_SUBSYS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 4 --warmup 3 --no-cfg4 --overlap-allreduce 0 > gpurun_out/bench_2gpu_serial.json 2> gpurun_out/bench_2gpu_serial.err
echo "bench2 serial exit=$?"; python - <<'PY'
import json
for f in ("bench_2gpu.json","bench_2gpu_serial.json"):
    try:
        line=[l for l in open("gpurun_out/"+f) if l.startswith("{")][-1]
        d=json.loads(line); print(f, d["value"], d["ms_per_step"], d["allreduce"])
    except Exception as e: print(f, "ERR", e)
PY
