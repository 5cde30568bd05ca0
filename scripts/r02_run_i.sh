#!/bin/bash
# full GPU suite + A/B of the grouped GEMM with per-tile field copies (new) against the previous build (libra_b200/build/alt.so)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_i.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_i.log)"; grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_i.log | head -10
bash scripts/ab_lib.sh libra_b200/build/alt.so --steps 4 --warmup 2
