#!/bin/bash
# final validation of the round: smoke, the default bench (both arms), launch list of an 8-layer step
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/smoke.log | cut -c1-200)"
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['share_of_step'], d['attention_roofline']['frac'], d['attention_roofline']['bwd_ms'], d['attention_roofline']['bwd_achieved_tflops'], d['cfg4']['value'], d['optimizer_ms'], d['gpu_eager_baseline']['value'], d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit=$?"; cut -c1-400 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step8_final.csv \
  python bench.py --layers 8 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cfg4 > gpurun_out/ncu_step8_final.log 2>&1
echo "ncu exit=$?"
