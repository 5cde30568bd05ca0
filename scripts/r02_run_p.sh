#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_p.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_p.log)"; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/pytest_p.log | head -8
timeout 300 python scripts/preprocess_bench.py --reps 10 > gpurun_out/preprocess_bench_words.jsonl 2>/dev/null; python - <<'PY'
import json
for l in open('gpurun_out/preprocess_bench_words.jsonl'):
    d=json.loads(l); print(d['workload'], round(d['cuda_ms_resident'],3), round(d['achieved_gbs_resident']), d.get('equal_to_cpu_reference'))
PY
LB_PREPROC_BYTES=1 timeout 300 python scripts/preprocess_bench.py --reps 10 > gpurun_out/preprocess_bench_bytes.jsonl 2>/dev/null; python - <<'PY'
import json
for l in open('gpurun_out/preprocess_bench_bytes.jsonl'):
    d=json.loads(l); print('bytes:', d['workload'], round(d['cuda_ms_resident'],3), round(d['achieved_gbs_resident']), d.get('equal_to_cpu_reference'))
PY
