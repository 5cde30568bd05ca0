"""Four representative grouped-GEMM launches for an `ncu --set full` capture (one warm-up each, then the captured ones):
  1. language q/k/v-shaped NT product 5880 x 4096 x 4096      2. the 9-problem q/k/v fan-out of a decoder layer
  3. gate|up with the SwiGLU epilogue + the two vision chains   4. weight gradient 4096 x 4096 x 5880 (both operands transposed)
  ncu --set full --clock-control none --import-source on -k regex:gemm_grouped -s 4 -c 4 -o gpurun_out/gemm python scripts/gemm_ncu.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libra_b200 import ops  # noqa: E402

dev, BF16 = "cuda", torch.bfloat16
rnd = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).to(BF16)
nl, nv, H, I = 5880, 2312, 4096, 11008
x = rnd(nl + nv, H)
Wq, Wk, Wv = (rnd(H, H, sc=0.02) for _ in range(3))
Aq, Ak, Av = (rnd(1024, H, sc=0.02) for _ in range(3))
Bq, Bk, Bv = (rnd(H, 1024, sc=0.02) for _ in range(3))
q, k, v = (torch.empty(nl + nv, H, device=dev, dtype=BF16) for _ in range(3))
mq, mk, mv = (torch.empty(nv, 1024, device=dev, dtype=BF16) for _ in range(3))
Wg, Wu = rnd(I, H, sc=0.02), rnd(I, H, sc=0.02)
Ag, Au = rnd(2752, H, sc=0.02), rnd(2752, H, sc=0.02)
Bg, Bu = rnd(I, 2752, sc=0.02), rnd(I, 2752, sc=0.02)
g, u, h = (torch.empty(nl + nv, I, device=dev, dtype=BF16) for _ in range(3))
mg, mu = (torch.empty(nv, 2752, device=dev, dtype=BF16) for _ in range(2))
dW = torch.empty(H, H, device=dev, dtype=BF16)
G = ops.gp


def cases():
    ops.gemm_grouped([G(x[:nl], Wq, q[:nl])])
    ops.gemm_grouped([G(x[nl:], Aq, mq), G(x[nl:], Ak, mk), G(x[nl:], Av, mv),
                      G(x[:nl], Wq, q[:nl]), G(x[:nl], Wk, k[:nl]), G(x[:nl], Wv, v[:nl]),
                      G(mq, Bq, q[nl:], wait_on=0), G(mk, Bk, k[nl:], wait_on=1), G(mv, Bv, v[nl:], wait_on=2)])
    ops.gemm_grouped([G(x[nl:], Ag, mg), G(x[nl:], Au, mu),
                      G(x[:nl], Wg, h[:nl], b2=Wu, epi=ops.EPI_SWIGLU, g=g[:nl], u=u[:nl]),
                      G(mg, Bg, g[nl:], wait_on=0), G(mu, Bu, u[nl:], wait_on=1)])
    ops.gemm_grouped([G(q[:nl], x[:nl], dW, ta=True, tb=True)])


cases()
torch.cuda.synchronize()
cases()
torch.cuda.synchronize()
print("done")
