#!/usr/bin/env python
"""Timeline of CTA 0 of the paired-tile attention forward (diagnostics): clock64 stamps per kv tile."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libra_b200 import _lib, ops, schedule

B, T, H, D = 4, 2048, 32, 128
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda: torch.randn(B * T, H * D, device=dev, generator=g).bfloat16()
Q, K0, V0, K1, V1 = mk(), mk(), mk(), mk(), mk()
flag = torch.zeros(B, T, dtype=torch.bool)
flag[:, 1:579] = True
w = schedule.build_attn_work(flag, B, T, True, dev)
qf = flag.reshape(-1).to(torch.uint8).to(dev)
trace = torch.zeros(64, 16, dtype=torch.int64, device=dev)
run = lambda: ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q2, None, None, None, B, T, H, D, True, 1 / math.sqrt(D), paired=True)
for _ in range(3):
    run()
_lib.call("lb_attn_fwd_pair_set_trace", ctypes.c_void_p(trace.data_ptr()))
run()
torch.cuda.synchronize()
_lib.call("lb_attn_fwd_pair_set_trace", None)
t = trace.cpu()
print("work item 0:", w.work_q2[0].tolist())
names = ["mma:P_A", "mma:A issued", "mma:P_B", "mma:B issued", "A:S seen", "A:max", "A:P arr", "B:S seen", "B:max", "B:P arr"]
t0 = int(t[0, 4])
for it in range(20):
    if int(t[it, 4]) == 0:
        break
    print(f"tile {it:2d}: " + "  ".join(f"{n}={int(t[it, s]) - t0}" for s, n in enumerate(names)))
