#!/bin/bash
# Round-2 validation run: GPU tests, bench, launch list of an 8-layer step, kernel micro-benches.  Logs -> gpurun_out/
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -c 600 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step8.csv \
  python bench.py --layers 8 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cfg4 --no-tokenizer > gpurun_out/ncu_step8.log 2>&1
echo "ncu exit=$?"
timeout 300 python scripts/attn_bench.py > gpurun_out/attn_bench.log 2>&1; tail -n 12 gpurun_out/attn_bench.log
timeout 400 python scripts/bench_generate.py > gpurun_out/generate.log 2>&1; tail -n 2 gpurun_out/generate.log
timeout 300 python scripts/decode_profile.py > gpurun_out/decode_profile.log 2>&1; tail -n 30 gpurun_out/decode_profile.log
