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
# kernels to time: argv (default: the product kernels).  An experimental kernel goes in its own process.
which = sys.argv[1:] or ["single", "stream", "dq", "dkv", "dqs", "dkvs"]
res = {}
if "single" in which:
    res["fwd"] = timeit(lambda: ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale, out=o))
if "stream" in which:
    o_ref, lse_ref = o.clone(), lse.clone()
    o2, lse2 = ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale, kernel="stream", plan=PLAN)
    torch.cuda.synchronize()
    print("stream vs single: max|dO| %.3e  max|dlse| %.3e" % ((o2.float() - o_ref.float()).abs().max().item(), (lse2 - lse_ref).abs().max().item()))
    res["fwd-stream"] = timeit(lambda: ops.attn_fwd(Q, K0, V0, K1, V1, qf, w.work_q, None, None, None, B, T, H, D, True, scale, out=o, kernel="stream", plan=PLAN))
for k, t in res.items():
    print(f"{k} {t*1e3:.0f} us = {fl/t/1e9:.0f} TF/s (algorithmic, causal)")
if "dq" in which and "dkv" in which:
    t_q = timeit(lambda: ops.attn_bwd_dq(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.work_q, None, None, B, T, H, D, True, scale))
    t_k = timeit(lambda: ops.attn_bwd_dkv(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.qtile_has, w.work_kv, None, None, B, T, H, D, True,
                                          scale, kv_cover=(True, True)))
    print(f"dq {t_q*1e3:.0f} us | dkv {t_k*1e3:.0f} us | bwd {2.5*fl/(t_q+t_k)/1e9:.0f} TF/s (algorithmic, causal)")
if "dqs" in which:
    dq_ref = ops.attn_bwd_dq(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.work_q, None, None, B, T, H, D, True, scale)
    dq_new = ops.attn_bwd_dq(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.work_q, None, None, B, T, H, D, True, scale, kernel="stream", plan=PLAN)
    torch.cuda.synchronize()
    d = (dq_new.float() - dq_ref.float()).abs()
    print("dq stream vs single: max|d| %.3e (max|ref| %.3e), bit-identical %s" % (d.max().item(), dq_ref.float().abs().max().item(), bool(torch.equal(dq_new, dq_ref))))
    t_s = timeit(lambda: ops.attn_bwd_dq(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.work_q, None, None, B, T, H, D, True, scale, kernel="stream", plan=PLAN))
    print(f"dq-stream {t_s*1e3:.0f} us = {1.5*fl/t_s/1e9:.0f} TF/s hardware (3 MMAs per tile)")
if "dkvs" in which:
    ref = ops.attn_bwd_dkv(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.qtile_has, w.work_kv, None, None, B, T, H, D, True, scale, kv_cover=(True, True))
    KPLAN = w.stream_plan(H, ops.sm_count(), ops.STREAM_HEAD_GROUP, float(os.environ.get("LB_PLAN_OVERHEAD", "2.0")), which="kv")
    print("dkv stream: supported, max items", ops.dkv_stream_limits(), "plan longest list", KPLAN[3])
    run = lambda: ops.attn_bwd_dkv(Q, K0, V0, K1, V1, dO, lse, delta, qf, w.qtile_has, w.work_kv, None, None, B, T, H, D, True, scale,
                                   kv_cover=(True, True), kernel="stream", plan=KPLAN)
    new = run()
    torch.cuda.synchronize()
    for nm, a, b_ in zip(("dK0", "dV0", "dK1", "dV1"), new, ref):
        print(f"dkv stream vs single {nm}: max|d| %.3e (max|ref| %.3e) identical %s" % ((a.float() - b_.float()).abs().max().item(), b_.float().abs().max().item(), bool(torch.equal(a, b_))))
    t_s = timeit(run)
    print(f"dkv-stream {t_s*1e3:.0f} us = {2.0*fl/t_s/1e9:.0f} TF/s hardware (4 MMAs per tile)")
if "fa2" in which:
    # informational: the flash-attn library, plain causal attention (use_bridge=False semantics)
    from libra_b200.utils.llama_flash_attn_monkey_patch import flash_attn_reference_point
    r = flash_attn_reference_point(B, T, H, D)
    if r is None or r[0] == "unavailable":
        print("fa2 unavailable:", r)
    else:
        print(f"fa2 (library, no bridge) fwd {r[0]*1e3:.0f} us = {fl/r[0]/1e9:.0f} TF/s | fwd+bwd {r[1]*1e3:.0f} us")
