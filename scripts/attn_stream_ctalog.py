#!/usr/bin/env python
"""Per-CTA log of the persistent attention forward (diagnostics): balance of the static item split, clk per tile."""
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
n_cta = torch.cuda.get_device_properties(0).multi_processor_count
log = torch.zeros(n_cta, 8, dtype=torch.int64, device=dev)
run = lambda: ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, 1 / math.sqrt(D), kernel="stream", plan=PLAN)
for _ in range(3):
    run()
_lib.call("lb_attn_fwd_stream_set_cta_log", ctypes.c_void_p(log.data_ptr()))
run()
torch.cuda.synchronize()
_lib.call("lb_attn_fwd_stream_set_cta_log", None)
t = log.cpu().double()
items, tiles, t_in, t_q, t_out = t[:, 1], t[:, 2], t[:, 3], t[:, 4], t[:, 5]
life = t_out - t_in
print(f"{n_cta} CTAs; items/CTA {items.min():.0f}..{items.max():.0f}; tiles/CTA min {tiles.min():.0f} mean {tiles.mean():.1f} max {tiles.max():.0f}")
print(f"CTA lifetime clk: min {life.min():.0f} mean {life.mean():.0f} max {life.max():.0f};  entry->first Q landed mean {(t_q - t_in).mean():.0f}")
print(f"clk per tile (lifetime/tiles): mean {(life / tiles).mean():.0f}  min {(life / tiles).min():.0f}  max {(life / tiles).max():.0f}")
