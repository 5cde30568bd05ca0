#!/usr/bin/env python
"""Where a skinny-GEMM launch spends its time: %globaltimer stamps per CTA (lb_gemm_skinny_set_trace) for the decode step's
shapes, L2 flushed before the traced launch.  Prints, per shape, the median over CTAs of each phase relative to the first CTA's
start, and the launch's span (first start -> last end).
    python scripts/gemm_skinny_trace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libra_b200 import _lib, ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.02).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M = 8
x4, x11 = rnd(M, 4096), rnd(M, 11008)
w = rnd(4096, 4096); c = torch.empty(M, 4096, device=dev, dtype=torch.bfloat16); res = rnd(M, 4096)
ws = [rnd(4096, 4096) for _ in range(3)]; cs = [torch.empty(M, 4096, device=dev, dtype=torch.bfloat16) for _ in range(3)]
wg, wu = rnd(11008, 4096), rnd(11008, 4096); act = torch.empty(M, 11008, device=dev, dtype=torch.bfloat16)
wd = rnd(4096, 11008)
t8 = rnd(M, 8); b8 = rnd(4096, 8)
cases = [("bridge 4096x8 (+addend)", lambda: ops.gemm_skinny([ops.gp(t8, b8, c, d=res)])),
         ("bridge 4096x8 (no addend)", lambda: ops.gemm_skinny([ops.gp(t8, b8, c)])),
         ("o-proj 4096x4096 (no residual)", lambda: ops.gemm_skinny([ops.gp(x4, w, c)])),
         ("o-proj 4096x4096 (+residual)", lambda: ops.gemm_skinny([ops.gp(x4, w, c, d=res)])),
         ("q|k|v one launch", lambda: ops.gemm_skinny([ops.gp(x4, wi, ci) for wi, ci in zip(ws, cs)])),
         ("gate|up SwiGLU", lambda: ops.gemm_skinny([ops.gp(x4, wg, act, epi=ops.EPI_SWIGLU, b2=wu)])),
         ("down 4096x11008 (+residual)", lambda: ops.gemm_skinny([ops.gp(x11, wd, c, d=res)]))]
names = ["start", "prologue done / W issued", "dependency resolved", "first stage landed", "accumulator complete",
         "partial published", "output written", "end", "epilogue: dependency resolved", "epilogue: accumulator in registers",
         "epilogue: first output element", "-"]
order = [0, 1, 2, 3, 4, 8, 9, 5, 10, 6, 7]
lib = _lib.load()
buf = torch.zeros(4096, 12, dtype=torch.int64, device=dev)
WARM = os.environ.get("WARM", "0") == "1"          # WARM=1: no L2 flush, the launch right after three identical ones
for name, fn in cases:
    for _ in range(3):
        fn()
    if not WARM:
        flush.zero_()
    buf.zero_()
    torch.cuda.synchronize()
    lib.lb_gemm_skinny_set_trace(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    lib.lb_gemm_skinny_set_trace(None)
    t = buf.cpu()
    used = t[:, 0] > 0
    t = t[used]
    t0 = int(t[:, 0].min())
    span = (int(t[:, 7].max()) - t0) / 1e3
    print(f"{name}: {int(used.sum())} CTAs, span {span:.2f} us, events {e0.elapsed_time(e1) * 1e3:.1f} us; last CTA start +{(int(t[:, 0].max()) - t0) / 1e3:.2f} us")
    for i in order:
        nm = names[i]
        col = t[:, i]
        col = col[col > 0]
        if len(col):
            rel = (col - t0).float() / 1e3
            print(f"    {nm:28s} median +{rel.median():6.2f} us   max +{rel.max():6.2f} us")
