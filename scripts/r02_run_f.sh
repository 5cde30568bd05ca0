#!/bin/bash
# Round-2 run f: full GPU suite after the ADVICE fixes, preprocessing bench + per-kernel times
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_f.log 2>&1
echo "pytest exit=$? $(tail -n 1 gpurun_out/pytest_f.log)"; grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_f.log | head -10
timeout 600 python scripts/preprocess_bench.py > gpurun_out/preprocess_bench.jsonl 2> gpurun_out/preprocess_bench.err; cut -c1-420 gpurun_out/preprocess_bench.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"resize_|tables_" --csv --log-file gpurun_out/preprocess_kernels.csv python scripts/preprocess_bench.py --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/preprocess_kernels.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki].split('(')[0]), {})[r[mi]]=r[vi]
for (i,k),m in list(d.items())[-9:]: print(i,k,m)
PY
