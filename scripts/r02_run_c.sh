#!/bin/bash
# Round-2 run c: skinny GEMM with three CTA slots per SM under the dependent-launch chain (A/B against two slots / serial launches)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_decode.py -m gpu -x -q -p no:cacheprovider --timeout 300 -k "skinny or pdl or decode or generate or dependent" > gpurun_out/pytest_c.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_c.log)"
LB_SKINNY_SLOTS=3 timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -p no:cacheprovider --timeout 300 -k "skinny or pdl" > gpurun_out/pytest_c3.log 2>&1
echo "pytest slots3 exit=$? $(tail -n 1 gpurun_out/pytest_c3.log)"
for cfg in "LB_PDL=0 LB_SKINNY_SLOTS=2" "LB_PDL=1 LB_SKINNY_SLOTS=2" "LB_PDL=1 LB_SKINNY_SLOTS=3" "LB_PDL=0 LB_SKINNY_SLOTS=3"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 python scripts/bench_generate.py > gpurun_out/generate_$tag.log 2>&1
  echo "$cfg: $(tail -n 1 gpurun_out/generate_$tag.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["decode_ms_per_step"], d["decode_frac_of_hbm_peak"], d["attn_decode"]["avg_launch_us"])')"
done
