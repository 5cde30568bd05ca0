#!/usr/bin/env python
"""Per-CTA log of the dK/dV kernel (diagnostics): where a CTA's time goes (prologue, loop, epilogue), clk per q tile."""
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
scale = 1 / math.sqrt(D)
o, lse = ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale)
_, delta = ops.attn_bwd_prepare(o, dO, None, B, T, H, D, want_dO_orig=False)
n_cta = w.work_kv.shape[0] * H
log = torch.zeros(n_cta, 8, dtype=torch.int64, device=dev)
run = lambda: ops.attn_bwd_dkv(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.qtile_has, w.work_kv, None, None, B, T, H, D, True, scale, kv_cover=(True, True))
for _ in range(3):
    run()
_lib.call("lb_attn_bwd_dkv_set_cta_log", ctypes.c_void_p(log.data_ptr()))
run()
torch.cuda.synchronize()
_lib.call("lb_attn_bwd_dkv_set_cta_log", None)
t = log.cpu().double()
t = t[t[:, 0] > 0]
n, t_in, t_list, t_kv, t_issued, t_done, t_out = (t[:, i] for i in range(7))
print(f"{len(t)} CTAs with tiles; q tiles per CTA min {n.min():.0f} mean {n.mean():.2f} max {n.max():.0f}")
print(f"mean clk: entry->tile list {(t_list - t_in).mean():.0f} | ->K/V landed {(t_kv - t_list).mean():.0f} | ->last MMA issued {(t_issued - t_kv).mean():.0f} "
      f"| ->all MMAs done {(t_done - t_issued).mean():.0f} | ->exit (epilogue) {(t_out - t_done).mean():.0f} | lifetime {(t_out - t_in).mean():.0f}")
loop = t_done - t_kv
for k in sorted(set(n.tolist())):
    m = n == k
    print(f"  {int(k):2d} q tiles: {int(m.sum()):5d} CTAs, loop clk/tile {(loop[m] / k).mean():.0f}, lifetime/tile {((t_out - t_in)[m] / k).mean():.0f}")
print(f"sum of lifetimes / (148 SMs) = {(t_out - t_in).sum() / 148:.0f} clk per SM")
