#!/usr/bin/env python
"""Micro-benchmark of the attention kernels alone (B=4, T=2048, H=32, D=128; the bench.py micro-batch shape).
Inputs (5 x 67 MB + outputs) exceed nothing special, so an L2 flush (256 MB write) precedes every timed launch."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libra_b200 import ops, schedule

B, T, H, D = 4, 2048, 32, 128
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda: torch.randn(B * T, H * D, device=dev, generator=g).bfloat16()
Q, K0, V0, K1, V1, dO = mk(), mk(), mk(), mk(), mk(), mk()
flag = torch.zeros(B, T, dtype=torch.bool)
flag[:, 1:579] = True
w = schedule.build_attn_work(flag, B, T, True, dev)
qf = flag.reshape(-1).to(torch.uint8).to(dev)
PLAN = None if os.environ.get("LB_STREAM_SNAKE") else w.stream_plan(H, ops.sm_count(), ops.STREAM_HEAD_GROUP, float(os.environ.get("LB_PLAN_OVERHEAD", "2.0")))
scale = 1 / math.sqrt(D)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
o, lse = ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale)
_, delta = ops.attn_bwd_prepare(o, dO, None, B, T, H, D, want_dO_orig=False)


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

fl = B * 4 * T * T * H * D / 2
t_f = timeit(lambda: ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale, out=o))
t_p = timeit(lambda: ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q2, None, None, None, B, T, H, D, True, scale, out=o, paired=True))
t_s = timeit(lambda: ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale, out=o, kernel="stream", plan=PLAN))
t_q = timeit(lambda: ops.attn_bwd_dq(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.work_q, None, None, B, T, H, D, True, scale))
t_k = timeit(lambda: ops.attn_bwd_dkv(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.qtile_has, w.work_kv, None, None, B, T, H, D, True,
                                      scale, kv_cover=(True, True)))
print(f"LB_EXP_POLY={os.environ.get('LB_EXP_POLY', 'default')}  fwd {t_f*1e3:.0f} us = {fl/t_f/1e9:.0f} TF/s | fwd-pair(poly={os.environ.get('LB_PAIR_EXP_POLY', '0')}) {t_p*1e3:.0f} us = {fl/t_p/1e9:.0f} TF/s | fwd-stream(poly={os.environ.get('LB_STREAM_EXP_POLY', '0')}) {t_s*1e3:.0f} us = {fl/t_s/1e9:.0f} TF/s | dq {t_q*1e3:.0f} us | "
      f"dkv {t_k*1e3:.0f} us | bwd {2.5*fl/(t_q+t_k)/1e9:.0f} TF/s (algorithmic, causal)")
