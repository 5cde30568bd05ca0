#!/usr/bin/env python
"""Summarise ncu output into markdown for profiles/.

  python scripts/ncu_summary.py launches <launches.csv> [title]       per-kernel totals and shares of a launch list
  python scripts/ncu_summary.py rep <file.ncu-rep> [regex]            key metrics per profiled launch (needs ncu here)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def launches(path, title="launch list"):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in r:
        if len(row) <= iv:
            continue
        v = float(row[iv].replace(",", ""))
        v = {"ns": v / 1e3, "ms": v * 1e3, "s": v * 1e6}.get(row[iu], v)
        name = re.sub(r"\(.*", "", row[ik])
        name = re.sub(r"^void ", "", name)[:80]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    n = sum(a[0] for a in agg.values())
    print(f"# {title}\n\n{n} launches, {tot / 1e3:.2f} ms summed device time (ncu: serialised, cold cache -- compare shares).\n")
    ours = sum(t for k, (c, t) in agg.items() if k.startswith("lb::"))
    gemm = sum(t for k, (c, t) in agg.items() if not k.startswith("lb::") and ("nvjet" in k or "cutlass" in k or "gemm" in k.lower() or "cublas" in k.lower()))
    print(f"libra_b200 kernels: {ours / tot * 100:.1f}% | cuBLAS GEMMs: {gemm / tot * 100:.1f}% | other torch kernels: {(tot - ours - gemm) / tot * 100:.1f}%\n")
    print("| share | total us | launches | avg us | kernel |\n|---:|---:|---:|---:|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        print(f"| {t / tot * 100:.2f}% | {t:.0f} | {c} | {t / c:.1f} | `{k}` |")


def rep(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full: {path}\n")
    seen = collections.Counter()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        if pattern and not re.search(pattern, name):
            continue
        seen[name] += 1
        if seen[name] > 2:
            continue
        print(f"## {name} (launch #{seen[name]})\n")
        for k, label in KEYS:
            if k in idx:
                print(f"- {label}: {r[idx[k]]} {units[idx[k]]}  (`{k}`)")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:])
    else:
        rep(*sys.argv[2:])
