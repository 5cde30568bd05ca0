#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
for sp in 0 2 3 4 6 10; do
  if [ $sp = 0 ]; then unset LB_DECODE_SPLIT; else export LB_DECODE_SPLIT=$sp; fi
  timeout 300 python scripts/bench_generate.py > gpurun_out/generate_split$sp.log 2>&1
  echo "split=$sp: $(tail -n 1 gpurun_out/generate_split$sp.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["decode_ms_per_step"], d["decode_frac_of_hbm_peak"], d["attn_decode"]["avg_launch_us"], d["attn_decode"]["frac"])')"
done
