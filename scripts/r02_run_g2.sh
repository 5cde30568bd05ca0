#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bandwidth.py tests/test_gpu_decode.py -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_g.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_g.log)"; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/pytest_g.log | head -10
for cfg in "LB_FOLD_DECODE_BRIDGE=0" "LB_FOLD_DECODE_BRIDGE=1" "LB_FOLD_DECODE_BRIDGE=0" "LB_FOLD_DECODE_BRIDGE=1"; do
  env $cfg timeout 300 python scripts/bench_generate.py > gpurun_out/generate_$cfg.log 2>&1
  echo "$cfg: $(tail -n 1 gpurun_out/generate_$cfg.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["decode_ms_per_step"], d["decode_frac_of_hbm_peak"], d["prefill_tokens_per_s"])')"
done
timeout 300 python scripts/decode_profile.py > gpurun_out/decode_profile_g.log 2>&1; sed -n 3,30p gpurun_out/decode_profile_g.log | cut -c1-130
