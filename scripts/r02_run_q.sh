#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out; rm -f gpurun_out/parity_tiny.jsonl
LB_PARITY_LOG=gpurun_out/parity_tiny.jsonl timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_q.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_q.log)"; cat gpurun_out/parity_tiny.jsonl
