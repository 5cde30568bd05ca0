#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_o.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_o.log)"; grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_o.log | head -10
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"swiglu|adamw" --csv --log-file gpurun_out/launches_sw.csv \
  python bench.py --layers 4 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cfg4 --no-tokenizer > gpurun_out/ncu_sw.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/launches_sw.csv x | sed -n 7,12p
