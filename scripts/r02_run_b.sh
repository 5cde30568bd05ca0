#!/bin/bash
# Round-2 run b: PDL decode chain + skinny-GEMM reduce fixes (tests, A/B), ncu of the streaming attention-backward kernels,
# GPU-busy trace of an 8-layer training step.  Logs -> gpurun_out/
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_decode.py -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_b.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_b.log)"
LB_PDL=0 timeout 300 python scripts/bench_generate.py > gpurun_out/generate_pdl0.log 2>&1; tail -n 1 gpurun_out/generate_pdl0.log | cut -c1-400
timeout 300 python scripts/bench_generate.py > gpurun_out/generate_pdl1.log 2>&1; tail -n 1 gpurun_out/generate_pdl1.log | cut -c1-400
timeout 300 python scripts/decode_profile.py > gpurun_out/decode_profile_b.log 2>&1; sed -n 3,30p gpurun_out/decode_profile_b.log | cut -c1-150
timeout 300 python scripts/attn_bench.py stream dqs dkvs > gpurun_out/attn_bench_b.log 2>&1; cat gpurun_out/attn_bench_b.log
for k in dq dkv; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_${k}_stream -s 2 -c 1 -f -o gpurun_out/attn_bwd_${k}_stream_r02 \
    python scripts/attn_bench.py ${k}s > gpurun_out/ncu_${k}.log 2>&1
  echo "ncu $k exit=$?"
done
timeout 600 python bench.py --layers 8 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cfg4 --busy-trace > gpurun_out/bench_busy8.json 2> gpurun_out/bench_busy8.err
echo "busy exit=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_busy8.json'))
b=d.get('busy'); print(d['ms_per_step'], {k:b[k] for k in b if k!='top'})
for r in b['top']: print(f"{r['ms']:9.3f} ms {r['launches']:5d}  {r['kernel'][:90]}")
PY
