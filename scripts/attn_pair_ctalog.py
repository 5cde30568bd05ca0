#!/usr/bin/env python
"""Per-CTA log of the paired-tile attention forward (diagnostics): where the SM time goes (prologue, loop, epilogue,
gaps between CTAs, tail)."""
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
n_cta = w.work_q2.shape[0] * H
log = torch.zeros(n_cta, 8, dtype=torch.int64, device=dev)
run = lambda: ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q2, None, None, None, B, T, H, D, True, 1 / math.sqrt(D), paired=True)
for _ in range(3):
    run()
_lib.call("lb_attn_fwd_pair_set_cta_log", ctypes.c_void_p(log.data_ptr()))
run()
torch.cuda.synchronize()
_lib.call("lb_attn_fwd_pair_set_cta_log", None)
t = log.cpu()
smid, sa, sb, t_in, t_q, t_mma, t_out = (t[:, i] for i in range(7))
steps = sa + sb
print(f"{n_cta} CTAs, {int(steps.sum())} lane-steps (64 keys x 128 rows); per SM:")
tot_busy = tot_span = 0
rows = []
for sm in sorted(set(smid.tolist())):
    idx = (smid == sm).nonzero().flatten()
    o = idx[t_in[idx].argsort()]
    span = int(t_out[o].max() - t_in[o].min())
    busy = int((t_out[o] - t_in[o]).sum())
    rows.append((sm, len(o), int(steps[o].sum()), span, busy))
spans = torch.tensor([r[3] for r in rows], dtype=torch.float64)
print(f"SMs {len(rows)}  span clk: min {spans.min():.0f} mean {spans.mean():.0f} max {spans.max():.0f}")
print(f"mean CTA-resident clk per SM {sum(r[4] for r in rows)/len(rows):.0f}   mean lane-steps per SM {sum(r[2] for r in rows)/len(rows):.1f}")
pro = (t_q - t_in).double(); loop = (t_mma - t_q).double(); epi = (t_out - t_mma).double()
print(f"per CTA: prologue(entry->Q landed) mean {pro.mean():.0f}  loop(Q->last MMA issued) mean {loop.mean():.0f}  "
      f"epilogue(last issue->exit) mean {epi.mean():.0f}")
both = (sa > 0) & (sb > 0)
print(f"loop clk per lane-step: two-lane CTAs {float(loop[both].sum() / steps[both].sum()):.0f}   single-lane CTAs "
      f"{float(loop[~both].sum() / max(1, int(steps[~both].sum()))):.0f}   ({int(both.sum())} / {int((~both).sum())} CTAs)")
for k in (1, 2, 4, 8, 16, 32):
    m = both & (torch.maximum(sa, sb) == k)
    if m.any():
        print(f"  two-lane CTAs with {k:2d} steps: n={int(m.sum()):4d}  prologue {pro[m].mean():.0f}  loop {loop[m].mean():.0f}  epilogue {epi[m].mean():.0f}")
