#!/usr/bin/env python
"""N1 micro-benchmark: the KV-cached decode attention kernel alone (lb_attn_decode) against the HBM roofline.
Algorithmic bytes per launch = B * kv_len * H*D * 2 (K and V) * 2 bytes.  An L2 flush precedes every timed launch."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libra_b200 import ops

dev = "cuda"
H, D = 32, 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peaks = json.load(open(_pk)) if os.path.exists(_pk) else {"hbm_gbs": 6650.0}
for B, T in ((8, 2048), (8, 4096), (32, 2048), (1, 4096)):
    g = torch.Generator(device=dev).manual_seed(0)
    C = H * D
    mk = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
    q = mk(B, C)
    Kfl, Vfl = mk(B, T, C), mk(B, T, C)
    out = torch.empty(B, C, dtype=torch.bfloat16, device=dev)
    run = lambda: ops.attn_decode(q, Kfl, Vfl, None, None, None, None, None, None, B, H, D, T, 1 / math.sqrt(D), out=out)
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    t = ts[len(ts) // 2]
    by = B * T * C * 2 * 2
    print(json.dumps({"kernel": "attn_decode_kernel<128> + combine", "B": B, "kv_len": T, "us": round(t * 1e3, 1),
                      "achieved_gbs": round(by / t / 1e6, 1), "peak_gbs": peaks["hbm_gbs"], "frac": round(by / t / 1e6 / peaks["hbm_gbs"], 3)}))
