#!/usr/bin/env python
"""Timeline of CTA 0 of the persistent dQ kernel (diagnostics): clock64 stamps per 128x64 tile."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libra_b200 import _lib, ops, schedule

B, T, H, D = 4, 2048, 32, 128
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda: torch.randn(B * T, H * D, device=dev, generator=g).bfloat16()
Q, K0, V0, K1, V1, dO = mk(), mk(), mk(), mk(), mk(), mk()
flag = torch.zeros(B, T, dtype=torch.bool)
flag[:, 1:579] = True
w = schedule.build_attn_work(flag, B, T, True, dev)
qf = flag.reshape(-1).to(torch.uint8).to(dev)
PLAN = w.stream_plan(H, ops.sm_count(), ops.STREAM_HEAD_GROUP)
scale = 1 / math.sqrt(D)
o, lse = ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale)
_, delta = ops.attn_bwd_prepare(o, dO, None, B, T, H, D, want_dO_orig=False)
trace = torch.zeros(64, 32, dtype=torch.int64, device=dev)
run = lambda: ops.attn_bwd_dq(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.work_q, None, None, B, T, H, D, True, scale, kernel="stream", plan=PLAN)
for _ in range(3):
    run()
_lib.call("lb_attn_bwd_dq_stream_set_trace", ctypes.c_void_p(trace.data_ptr()))
run()
torch.cuda.synchronize()
_lib.call("lb_attn_bwd_dq_stream_set_trace", None)
t = trace.cpu()
names = {0: "a:start", 1: "a:Sfree", 2: "a:QK ok", 3: "a:dPfree", 4: "a:commit", 5: "b:start", 6: "b:dQfree", 7: "b:dS seen", 8: "b:commit",
         10: "c0:wait", 11: "c0:seen", 12: "c0:loaded", 13: "c0:computed", 14: "c0:arrived", 15: "c1:wait", 16: "c1:seen", 17: "c1:loaded", 18: "c1:computed", 19: "c1:arrived"}
t0 = int(t[0, 0])
for it in range(48):
    if int(t[it, 0]) == 0:
        break
    print(f"tile {it:2d}: " + " ".join(f"{n}={int(t[it, s]) - t0 if int(t[it, s]) else -1}" for s, n in names.items()))
for it in range(8):
    if int(t[it, 21]) == 0:
        continue
    print(f"item {it}: " + " ".join(f"{n}={int(t[it, 20 + s]) - t0}" for s, n in enumerate(["e:start", "e:dQfull", "e:released", "e:stored"])))
