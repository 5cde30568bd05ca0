#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
for i in 1 2; do for cfg in "LB_SKINNY_FULLTOK=1" "LB_SKINNY_FULLTOK=0"; do
  env $cfg timeout 300 python scripts/bench_generate.py > gpurun_out/generate_$cfg.log 2>&1
  echo "$cfg: $(tail -n 1 gpurun_out/generate_$cfg.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["decode_ms_per_step"], d["decode_frac_of_hbm_peak"], d["prefill_tokens_per_s"])')"
done; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
