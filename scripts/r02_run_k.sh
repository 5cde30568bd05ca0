#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bandwidth.py tests/test_gpu_decode.py tests/test_gpu_model.py -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_k.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_k.log)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attn_prep|rmsnorm_fwd" --csv --log-file gpurun_out/launches_prep.csv \
  python bench.py --layers 4 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cfg4 --no-tokenizer > gpurun_out/ncu_prep.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/launches_prep.csv x | head -14
