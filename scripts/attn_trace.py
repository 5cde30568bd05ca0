#!/usr/bin/env python
"""Timeline of one attention-forward CTA (diagnostics): clock64 stamps per kv tile of CTA (0,0)."""
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
trace = torch.zeros(64, 8, dtype=torch.int64, device=dev)
for _ in range(3):
    ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, 1 / math.sqrt(D))
_lib.call("lb_attn_fwd_set_trace", ctypes.c_void_p(trace.data_ptr()))
ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, 1 / math.sqrt(D))
torch.cuda.synchronize()
_lib.call("lb_attn_fwd_set_trace", None)
t = trace.cpu()
print("work item 0:", w.work_q[0].tolist())
names = ["mma:K ready", "mma:QK issued", "mma:P seen", "mma:PV issued", "sm:S seen", "sm:max done", "sm:exchanged", "sm:P arrived"]
t0 = int(t[0, 0])
for it in range(20):
    if int(t[it, 0]) == 0:
        break
    row = [int(t[it, s]) - t0 for s in range(8)]
    print(f"tile {it:2d}: " + "  ".join(f"{n}={v}" for n, v in zip(names, row)))
