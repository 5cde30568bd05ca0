#!/bin/bash
# A/B two builds of liblibra_b200.so on the same box: bash scripts/ab_lib.sh <alt.so> [bench args...]
# Runs new, alt, new, alt (interleaved, so clock/power drift shows up as disagreement between the repeats).
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
alt=$1; shift
mkdir -p gpurun_out
cp libra_b200/liblibra_b200.so /tmp/lb_new.so
cp "$alt" /tmp/lb_alt.so
for i in 1 2; do
  for v in new alt; do
    cp /tmp/lb_$v.so libra_b200/liblibra_b200.so
    timeout 300 python bench.py --no-cpu-baseline --no-gpu-baseline --no-cfg4 --no-e2e --steps 5 --warmup 3 "$@" > gpurun_out/ab_${v}_$i.json 2> gpurun_out/ab_${v}_$i.err
    python - "$v$i" gpurun_out/ab_${v}_$i.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d["roofline"]; a = d["attention_roofline"]
    print(sys.argv[1], "ms/step %.1f" % d["ms_per_step"], "gemm avg ms %.4f frac %.3f" % (r["avg_launch_ms"], r["frac"]), "opt ms %.1f" % d["optimizer_ms"],
          "attn fwd %.4f bwd %s" % (a["avg_launch_ms"], a.get("bwd_ms")), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "failed", e)
P
  done
done
cp /tmp/lb_new.so libra_b200/liblibra_b200.so
