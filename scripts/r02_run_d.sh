#!/bin/bash
# Round-2 run d: CLIP preprocessing kernel (tests + bench), sampling generate tests
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_decode.py -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_d.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_d.log)"; grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/pytest_d.log | head -20
timeout 600 python scripts/preprocess_bench.py > gpurun_out/preprocess_bench.jsonl 2> gpurun_out/preprocess_bench.err; cat gpurun_out/preprocess_bench.jsonl; tail -3 gpurun_out/preprocess_bench.err
timeout 600 python bench.py --layers 8 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cfg4 --busy-trace > gpurun_out/bench_busy8b.json 2> gpurun_out/bench_busy8b.err
echo "busy exit=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_busy8b.json'))
b=d.get('busy'); print(d['ms_per_step'], {k:b[k] for k in b if k not in('top','largest_gaps')})
for r in b['largest_gaps']: print(r)
PY
