#!/usr/bin/env python
"""Where does the persistent forward differ from the baseline kernel?  (diagnostics)"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libra_b200 import ops, schedule

B, T, H, D = 4, 2048, 32, 128
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda: torch.randn(B * T, H * D, device=dev, generator=g).bfloat16()
Q, K0, V0, K1, V1 = mk(), mk(), mk(), mk(), mk()
flag = torch.zeros(B, T, dtype=torch.bool)
flag[:, 1:579] = True
w = schedule.build_attn_work(flag, B, T, True, dev)
qf = flag.reshape(-1).to(torch.uint8).to(dev)
plan = None if os.environ.get("LB_STREAM_SNAKE") else w.stream_plan(H, ops.sm_count(), ops.STREAM_HEAD_GROUP)
scale = 1 / math.sqrt(D)
o, lse = ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale)
for rep in range(3):
    o2, lse2 = ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale, kernel="stream", plan=plan)
    torch.cuda.synchronize()
    d = (o2.float() - o.float()).abs().view(B, T // 128, 128, H, D).amax(dim=(2, 4))      # [B, q_tile, H]
    bad = (d > 0.05).nonzero().tolist()
    print(f"rep {rep}: max diff {d.max().item():.3e}, bad (b, q_tile, head) count {len(bad)} of {d.numel()}; lse max diff {(lse2 - lse).abs().max().item():.3e}")
    print("  first:", bad[:24])
    if bad:
        b, qt, h = bad[0]
        rows = (o2.float() - o.float()).abs().view(B, T, H, D)[b, qt * 128:(qt + 1) * 128, h].amax(dim=1)
        print("  rows of the first bad tile with diff > 0.05:", (rows > 0.05).nonzero().flatten().tolist()[:40], " n =", int((rows > 0.05).sum()))
        cols = (o2.float() - o.float()).abs().view(B, T, H, D)[b, qt * 128:(qt + 1) * 128, h].amax(dim=0)
        print("  cols:", (cols > 0.05).nonzero().flatten().tolist()[:40], " n =", int((cols > 0.05).sum()))
