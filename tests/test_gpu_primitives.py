"""tcgen05 building blocks on the real device: the four UMMA operand layouts through lb_gemm_bf16, and the
TMEM-A (TS) form the attention kernels use for P.V."""
import pytest
import torch

from gpu_util import need_gpu, assert_close, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("K", [64, 128])
def test_probe_ts(mode, K):
    need_gpu()
    from libra_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(K + mode)
    a = torch.randn(128, K, device="cuda", generator=g).bfloat16()
    bt = torch.randn(128, K, device="cuda", generator=g).bfloat16()          # logical B^T: [N, K]
    b = bt.contiguous() if mode == 0 else bt.t().contiguous()               # stored [N,K] or [K,N]
    d = ops.probe_umma(mode, a, b)
    torch.cuda.synchronize()
    want = a.float() @ bt.float().t()
    assert_close(d, want, rtol=1e-3, atol=1e-2, msg=f"probe mode={mode} K={K}")


@pytest.mark.parametrize("trans_a,trans_b", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (256, 384, 192), (300, 200, 136), (4096, 1024, 1024)])
def test_gemm_layouts(trans_a, trans_b, M, N, K):
    need_gpu()
    from libra_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    Bm = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    want = A.float() @ Bm.float().t()
    a = A.t().contiguous() if trans_a else A
    b = Bm.t().contiguous() if trans_b else Bm
    if (trans_a and M % 8) or (trans_b and N % 8):
        pytest.skip("leading dimension must be a multiple of 8")
    c = ops.gemm(a, b, trans_a=trans_a, trans_b=trans_b, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_err(c, want) < 2e-3, rel_err(c, want)
    assert_close(c, want, rtol=2e-3, atol=0.05 * K ** 0.5 / 8)


def test_gemm_epilogues():
    need_gpu()
    from libra_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(384, 256, device="cuda", generator=g).bfloat16()
    W = (torch.randn(520, 256, device="cuda", generator=g) * 0.1).bfloat16()
    bias = torch.randn(520, device="cuda", generator=g).bfloat16()
    z = A.float() @ W.float().t() + bias.float()
    c = ops.gemm(A, W, bias=bias, act=1)
    assert_close(c, z * torch.sigmoid(1.702 * z), rtol=2e-2, atol=2e-2)
    # accumulate into fp32 and bf16
    base = torch.randn(384, 520, device="cuda", generator=g)
    c32 = base.clone()
    ops.gemm(A, W, out=c32, accumulate=True)
    assert_close(c32, base + A.float() @ W.float().t(), rtol=2e-3, atol=2e-2)
    # ragged N (514, the vision head width) and bf16 out
    W2 = W[:514].contiguous()
    c2 = ops.gemm(A, W2)
    assert c2.shape == (384, 514)
    assert_close(c2, A.float() @ W2.float().t(), rtol=2e-2, atol=2e-2)
