#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_wrapper.py tests/test_gpu_vision.py -m gpu -x -q -p no:cacheprovider --timeout 90 > gpurun_out/pytest_r.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_r.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r.log | head -5
