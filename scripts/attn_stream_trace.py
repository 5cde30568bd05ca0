#!/usr/bin/env python
"""Timeline of CTA 0 of the persistent attention forward (diagnostics): clock64 stamps per kv tile."""
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
PLAN = None if os.environ.get("LB_STREAM_SNAKE") else w.stream_plan(H, ops.sm_count(), ops.STREAM_HEAD_GROUP, float(os.environ.get("LB_PLAN_OVERHEAD", "2.0")))
trace = torch.zeros(64, 32, dtype=torch.int64, device=dev)
run = lambda: ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, 1 / math.sqrt(D), kernel="stream", plan=PLAN)
for _ in range(3):
    run()
_lib.call("lb_attn_fwd_stream_set_trace", ctypes.c_void_p(trace.data_ptr()))
run()
torch.cuda.synchronize()
_lib.call("lb_attn_fwd_stream_set_trace", None)
t = trace.cpu()
names = ["q:waitS", "q:Sfree", "q:KQ ok", "q:issued", "q:commit", "p:wait", "p:VO ok", "p:P seen", "p:issued", "p:commit",
         "s:waitS", "s:S seen", "s:loaded", "s:max", "s:prevmax", "s:pub", "s:Pfree", "s:P arr"]
t0 = int(t[0, 11])
for it in range(48):
    if int(t[it, 11]) == 0:
        break
    print(f"tile {it:2d} (wg {it & 1}): " + " ".join(f"{n}={int(t[it, s]) - t0 if int(t[it, s]) else -1}" for s, n in enumerate(names)))
enames = ["e:start", "e:dep", "e:Ofinal", "e:Ofree", "e:stored"]
for it in range(8):
    if int(t[it, 19]) == 0:
        break
    print(f"item {it}: " + " ".join(f"{n}={int(t[it, 18 + s]) - t0}" for s, n in enumerate(enames)))
