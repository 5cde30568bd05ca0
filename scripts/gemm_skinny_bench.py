#!/usr/bin/env python
"""Weight-streaming rate of the skinny GEMM (csrc/gemm_skinny.cu) on the decode step's shapes, L2 flushed before every launch.
    LB_SKINNY_UNITS_X2=4 python scripts/gemm_skinny_bench.py"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libra_b200 import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.02).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6550.0


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


M = int(os.environ.get("M", "8"))
x4, x11 = rnd(M, 4096), rnd(M, 11008)
cases = []
w = rnd(4096, 4096); c = torch.empty(M, 4096, device=dev, dtype=torch.bfloat16)
cases.append(("o-proj 4096x4096 (+residual)", lambda: ops.gemm_skinny([ops.gp(x4, w, c, d=c)]), w.numel() * 2))
ws = [rnd(4096, 4096) for _ in range(3)]; cs = [torch.empty(M, 4096, device=dev, dtype=torch.bfloat16) for _ in range(3)]
a8 = [rnd(8, 4096) for _ in range(2)]; c8 = [torch.empty(M, 8, device=dev, dtype=torch.bfloat16) for _ in range(2)]
cases.append(("q|k|v + 2 bridge downs, one launch", lambda: ops.gemm_skinny([ops.gp(x4, wi, ci) for wi, ci in zip(ws, cs)] + [ops.gp(x4, wi, ci) for wi, ci in zip(a8, c8)]), 3 * w.numel() * 2))
wg, wu = rnd(11008, 4096), rnd(11008, 4096); act = torch.empty(M, 11008, device=dev, dtype=torch.bfloat16)
cases.append(("gate|up SwiGLU 2x11008x4096", lambda: ops.gemm_skinny([ops.gp(x4, wg, act, epi=ops.EPI_SWIGLU, b2=wu)]), 2 * wg.numel() * 2))
wd = rnd(4096, 11008)
cases.append(("down 4096x11008 (+residual)", lambda: ops.gemm_skinny([ops.gp(x11, wd, c, d=c)]), wd.numel() * 2))
wl = rnd(32000, 4096); cl = torch.empty(M, 32000, device=dev, dtype=torch.bfloat16)
cases.append(("lm_head 32000x4096", lambda: ops.gemm_skinny([ops.gp(x4, wl, cl)]), wl.numel() * 2))
tot_us = tot_b = 0
for name, fn, nbytes in cases:
    t = timeit(fn)
    print(f"{name:40s} {t * 1e3:7.1f} us  {nbytes / t / 1e6:7.0f} GB/s = {nbytes / t / 1e6 / peak:.2f} of the measured HBM peak ({peak:.0f})")
