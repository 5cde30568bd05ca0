"""HBM-bound kernels vs the oracle / plain fp32 torch on the same inputs (bf16 I/O => bf16 tolerances;
integer results bit exact)."""
import pytest
import torch

from gpu_util import need_gpu, assert_close, rel_err
from oracle import libra_oracle as O

pytestmark = pytest.mark.gpu
dev = "cuda"


def _gen(seed):
    return torch.Generator(device=dev).manual_seed(seed)


@pytest.mark.parametrize("rows,cols", [(37, 256), (611, 4096), (100, 6144), (5, 64)])
def test_rmsnorm_routed_fwd_bwd(rows, cols):
    need_gpu()
    from libra_b200 import ops
    g = _gen(rows + cols)
    x = (torch.randn(rows, cols, device=dev, generator=g) * 2).bfloat16()
    wl = (1 + 0.1 * torch.randn(cols, device=dev, generator=g)).bfloat16()
    wv = (1 + 0.1 * torch.randn(cols, device=dev, generator=g)).bfloat16()
    flag = (torch.rand(rows, device=dev, generator=g) > 0.6)
    y, rstd = ops.rmsnorm_fwd(x, wl, wv, flag.to(torch.uint8), 1e-6)
    xf = x.float().requires_grad_(True)
    wlf, wvf = wl.float().requires_grad_(True), wv.float().requires_grad_(True)
    w_rows = torch.where(flag[:, None], wvf, wlf)
    want = w_rows * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))
    assert_close(y, want, rtol=1e-2, atol=1e-2)
    # exactness vs the oracle's bf16 path (same cast points): allow 1 bf16 ulp
    yo = torch.where(flag[:, None], O.rmsnorm(x, wv, 1e-6), O.rmsnorm(x, wl, 1e-6))
    assert ((y.float() - yo.float()).abs() <= 0.008 * yo.float().abs() + 1e-6).all()
    dy = torch.randn(rows, cols, device=dev, generator=g).bfloat16()
    res = torch.randn(rows, cols, device=dev, generator=g).bfloat16()
    dx, dwl, dwv = ops.rmsnorm_bwd(dy, x, wl, wv, flag.to(torch.uint8), rstd, residual_grad=res)
    want.backward(dy.float())
    assert_close(dx, xf.grad + res.float(), rtol=2e-2, atol=2e-2)
    assert rel_err(dwl, wlf.grad) < 1e-3 and rel_err(dwv, wvf.grad) < 1e-3


@pytest.mark.parametrize("rows,cols", [(577, 1024), (33, 128)])
def test_layernorm_fwd_bwd(rows, cols):
    need_gpu()
    from libra_b200 import ops
    g = _gen(rows)
    x = (torch.randn(rows, cols, device=dev, generator=g) * 2 + 0.5).bfloat16()
    w = (1 + 0.1 * torch.randn(cols, device=dev, generator=g)).bfloat16()
    b = (0.1 * torch.randn(cols, device=dev, generator=g)).bfloat16()
    y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5)
    xf, wf, bf = x.float().requires_grad_(True), w.float().requires_grad_(True), b.float().requires_grad_(True)
    want = torch.nn.functional.layer_norm(xf, (cols,), wf, bf, 1e-5)
    assert_close(y, want, rtol=1e-2, atol=1e-2)
    dy = torch.randn(rows, cols, device=dev, generator=g).bfloat16()
    dx, dw, db = ops.layernorm_bwd(dy, x, w, mean, rstd)
    want.backward(dy.float())
    assert_close(dx, xf.grad, rtol=2e-2, atol=2e-2)
    assert rel_err(dw, wf.grad) < 1e-3 and rel_err(db, bf.grad) < 1e-3


def test_swiglu_and_quick_gelu():
    need_gpu()
    from libra_b200 import ops
    g = _gen(3)
    gu = torch.randn(300, 2 * 352, device=dev, generator=g).bfloat16()
    gate, up = gu[:, :352], gu[:, 352:]          # strided views (fused gate|up buffer)
    out = ops.swiglu_fwd(gate, up)
    gf, uf = gate.float().requires_grad_(True), up.float().requires_grad_(True)
    want = torch.nn.functional.silu(gf) * uf
    assert_close(out, want, rtol=1e-2, atol=1e-2)
    d = torch.randn(300, 352, device=dev, generator=g).bfloat16()
    dg, du = ops.swiglu_bwd(d, gate, up)
    want.backward(d.float())
    assert_close(dg, gf.grad, rtol=2e-2, atol=2e-2)
    assert_close(du, uf.grad, rtol=2e-2, atol=2e-2)
    x = torch.randn(77, 256, device=dev, generator=g).bfloat16()
    bias = torch.randn(256, device=dev, generator=g).bfloat16()
    y = ops.bias_quick_gelu_fwd(x, bias)
    z = (x.float() + bias.float()).bfloat16().float().requires_grad_(True)
    wq = O.quick_gelu(z)
    assert_close(y, wq, rtol=1e-2, atol=1e-2)
    dy = torch.randn(77, 256, device=dev, generator=g).bfloat16()
    wq.backward(dy.float())
    assert_close(ops.bias_quick_gelu_bwd(dy, x, bias), z.grad, rtol=2e-2, atol=2e-2)


def test_gather_and_embeddings():
    need_gpu()
    from libra_b200 import ops
    g = _gen(4)
    src = torch.randn(50, 128, device=dev, generator=g).bfloat16()
    idx = torch.randperm(50, device=dev, generator=g).to(torch.int32)
    assert torch.equal(ops.gather_rows(src, idx), src[idx.long()])
    table = torch.randn(320, 64, device=dev, generator=g).bfloat16()
    ids = torch.randint(0, 320, (91,), device=dev, generator=g)
    assert torch.equal(ops.embed_lang(ids, table), table[ids])
    t0 = torch.randn(514, 32, device=dev, generator=g).bfloat16()
    t1 = torch.randn(514, 32, device=dev, generator=g).bfloat16()
    i0 = torch.randint(0, 514, (40,), device=dev, generator=g)
    i1 = torch.randint(0, 514, (40,), device=dev, generator=g)
    sig = torch.randn(200, 48, device=dev, generator=g).bfloat16()
    srow = torch.randint(0, 200, (40,), device=dev, generator=g).to(torch.int32)
    cat = ops.embed_vision_cat(i0, i1, t0, t1, sig, srow, 48)
    assert torch.equal(cat, torch.cat([t0[i0], t1[i1], sig[srow.long()]], -1))
    cat0 = ops.embed_vision_cat(i0, i1, t0, t1, None, None, 48)
    assert torch.equal(cat0[:, 64:], torch.zeros(40, 48, device=dev, dtype=torch.bfloat16))
    dy = torch.randn(40, 112, device=dev, generator=g).bfloat16()
    dt = torch.zeros(514, 32, device=dev)
    ops.embed_bwd(i1, dy, 32, 32, dt)
    want = torch.zeros(514, 32, device=dev).index_add_(0, i1, dy[:, 32:64].float())
    assert_close(dt, want, rtol=1e-5, atol=1e-5)


def test_lfq_pack_bit_exact(golden):
    need_gpu()
    from libra_b200 import ops
    g = golden("lfq")
    x = g["x"].permute(0, 2, 3, 1).reshape(2 * 576, 18).contiguous().to(dev)      # [n_img*tokens, 18]
    for dt in (torch.float32, torch.bfloat16):
        ids = ops.lfq_pack(x.to(dt), 2, 576, 2, 9, 32000, 32512, 32513)
        want_idx = O.lfq_indices(x.to(dt).float().view(2, 576, 18).cpu(), 2, 9)
        want = O.image_token_ids(want_idx, 32000, 512)
        assert torch.equal(ids.cpu(), want)
        if dt == torch.float32:
            ref_idx = g["indices"].reshape(2, 576, 2).to(torch.int64)                  # produced by the reference LFQ
            assert torch.equal(ids[:, :, 1:-1].cpu().permute(1, 2, 0) - 32000, ref_idx)
    idx = (ids[:, :, 1:-1].permute(1, 2, 0) - 32000).contiguous()
    codes = ops.lfq_unpack(idx, 2, 9, torch.float32)
    assert torch.equal(codes.cpu(), O.lfq_codes(idx.cpu(), 9))
    # large random round trip: unpack(pack(x)) == sign(x)
    big = torch.randn(64 * 576, 18, device=dev)
    ids = ops.lfq_pack(big, 64, 576, 2, 9, 0, 512, 513)
    codes = ops.lfq_unpack((ids[:, :, 1:-1].permute(1, 2, 0)).contiguous(), 2, 9, torch.float32)
    assert torch.equal(codes.view(-1, 18), torch.where(big > 0, 1.0, -1.0))


@pytest.mark.parametrize("vocab,ld", [(32000, 32000), (514, 514), (514, 520), (834, 840)])
def test_cross_entropy(vocab, ld):
    need_gpu()
    from libra_b200 import ops
    g = _gen(vocab)
    rows = 97
    buf = torch.zeros(rows, ld, device=dev, dtype=torch.bfloat16)
    logits = (torch.randn(rows, vocab, device=dev, generator=g) * 3).bfloat16()
    buf[:, :vocab] = logits
    labels = torch.randint(0, vocab, (rows,), device=dev, generator=g)
    labels[::7] = -100
    lf = logits.float().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(lf, labels, ignore_index=-100, reduction="none")
    view = buf[:, :vocab] if ld != vocab else buf
    loss = ops.cross_entropy_fwd_bwd(view, labels, vocab, 0.5)
    assert_close(loss, want, rtol=1e-3, atol=2e-3)
    (want.sum() * 0.5).backward()
    assert_close(view, lf.grad, rtol=2e-2, atol=2e-3)


def test_attn_prep_fwd_bwd():
    """bridge add + RoPE + un-permute vs the oracle formulas, and its adjoint vs autograd."""
    need_gpu()
    from libra_b200 import ops, schedule
    g = _gen(9)
    B, T, H, D, R = 2, 50, 2, 128, 8
    C = H * D
    flag = torch.zeros(B, T, dtype=torch.bool, device=dev)
    flag[0, 1:30] = True
    flag[1, 10:41] = True
    rt = schedule.build_routing(flag)
    N = B * T
    mk = lambda *s: (torch.randn(*s, device=dev, generator=g)).bfloat16()
    q, k, v = mk(N, C), mk(N, C), mk(N, C)                      # sorted rows
    tk, tv = mk(N, R), mk(N, R)
    Bk_l, Bk_v, Bv_l, Bv_v = (mk(C, R) * 0.3 for _ in range(4))
    Bk_l, Bk_v, Bv_l, Bv_v = (t.bfloat16() for t in (Bk_l, Bk_v, Bv_l, Bv_v))
    pos = torch.arange(T, device=dev, dtype=torch.int32).repeat(B)
    cos, sin = O.rope_tables(D, 2048, 10000.0, torch.float32, dev)
    cos_t, sin_t = cos[:, :D // 2].contiguous(), sin[:, :D // 2].contiguous()
    nl = rt.n_lang
    kc = torch.cat([torch.addmm(k[:nl], tk[:nl], Bk_l.t()), torch.addmm(k[nl:], tk[nl:], Bk_v.t())])
    vc = torch.cat([torch.addmm(v[:nl], tv[:nl], Bv_l.t()), torch.addmm(v[nl:], tv[nl:], Bv_v.t())])
    Q, Kfv, Kfl, Vfv, Vfl = ops.attn_prep_fwd(q, k, kc, v, vc, rt.flag_sorted, rt.inv, pos, cos_t, sin_t, H, D)

    leaves = [t.float().requires_grad_(True) for t in (q, k, v, tk, tv)]
    qf, kf, vf, tkf, tvf = leaves
    fs = rt.flag_sorted.bool()[:, None]
    kb = torch.where(fs, tkf @ Bk_v.float().t(), tkf @ Bk_l.float().t())
    vb = torch.where(fs, tvf @ Bv_v.float().t(), tvf @ Bv_l.float().t())
    inv = rt.inv.long()

    def rope(x):    # x: [N(orig), C]
        xh = x.view(N, H, D)
        c = cos[pos.long()][:, None, :]
        s = sin[pos.long()][:, None, :]
        rot = torch.cat((-xh[..., D // 2:], xh[..., :D // 2]), -1)
        return (xh * c + rot * s).reshape(N, C)
    fo = rt.flag_orig.bool()[:, None]
    Qw = rope(qf[inv])
    Ks, Kc = rope(kf[inv]), rope((kf + kb)[inv])
    Kfv_w, Kfl_w = torch.where(fo, Ks, Kc), torch.where(fo, Kc, Ks)
    Vs, Vc = vf[inv], (vf + vb)[inv]
    Vfv_w, Vfl_w = torch.where(fo, Vs, Vc), torch.where(fo, Vc, Vs)
    for got, want, nm in ((Q, Qw, "Q"), (Kfv, Kfv_w, "Kfv"), (Kfl, Kfl_w, "Kfl"), (Vfv, Vfv_w, "Vfv"), (Vfl, Vfl_w, "Vfl")):
        assert_close(got, want, rtol=2e-2, atol=3e-2, msg=nm)
    # the decode step's form: the rank-8 bridge products computed inside the prologue (lb_attn_prep_fwd_bridge), also with the
    # K/V rows redirected (kv_row) as the KV cache does
    outs = ops.attn_prep_fwd_bridge(q, k, v, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v, rt.flag_sorted, rt.inv, pos, cos_t, sin_t, H, D)
    for got, ref, want, nm in zip(outs, (Q, Kfv, Kfl, Vfv, Vfl), (Qw, Kfv_w, Kfl_w, Vfv_w, Vfl_w), ("Q", "Kfv", "Kfl", "Vfv", "Vfl")):
        assert_close(got, want, rtol=2e-2, atol=3e-2, msg=nm + " (folded bridge)")
        d = (got.float() - ref.float()).abs()          # vs the addmm-produced variant: same rounding sequence, fp32 sum order aside
        assert float((d > 0).float().mean()) < 2e-2 and float(d.max()) <= 0.0625, nm
    rows = torch.randperm(N, device=dev, generator=g).to(torch.int32)
    big = [torch.zeros(N, C, dtype=torch.bfloat16, device=dev) for _ in range(4)]
    ops.attn_prep_fwd_bridge(q, k, v, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v, rt.flag_sorted, rt.inv, pos, cos_t, sin_t, H, D,
                             kv_out=big, kv_row=rows)
    for got, ref in zip(big, outs[1:]):
        assert torch.equal(got[rows.long()], ref)
    grads = [mk(N, C) for _ in range(5)]
    dq, dk, dv, dkb, dvb = ops.attn_prep_bwd(*grads, rt.flag_sorted, rt.inv, pos, cos_t, sin_t, H, D)
    kb.retain_grad(); vb.retain_grad()
    loss = sum((w * gr.float()).sum() for w, gr in zip((Qw, Kfv_w, Kfl_w, Vfv_w, Vfl_w), grads))
    loss.backward()
    assert_close(dq, qf.grad, rtol=2e-2, atol=3e-2, msg="dq")
    assert_close(dk, kf.grad, rtol=2e-2, atol=5e-2, msg="dk")
    assert_close(dv, vf.grad, rtol=2e-2, atol=5e-2, msg="dv")
    assert_close(dkb, kb.grad, rtol=2e-2, atol=3e-2, msg="dkb")
    assert_close(dvb, vb.grad, rtol=2e-2, atol=3e-2, msg="dvb")


def test_fused_adamw_matches_torch():
    need_gpu()
    from libra_b200.optim import FlatAdamW
    g = _gen(21)
    n = 8 * 1000 + 5
    p0 = torch.randn(n, device=dev, generator=g).bfloat16()
    ours_p, grad = p0.clone(), torch.zeros(n, device=dev, dtype=torch.bfloat16)
    opt = FlatAdamW(ours_p, grad, lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    ref_p = torch.nn.Parameter(p0.float().clone())
    ref = torch.optim.AdamW([ref_p], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    for _ in range(5):
        gr = torch.randn(n, device=dev, generator=g).bfloat16()
        grad.copy_(gr)
        ref_p.grad = gr.float()
        opt.step()
        ref.step()
    # bf16 parameter/state storage: compare against the fp32 optimizer within bf16 resolution of the values
    assert_close(ours_p, ref_p.detach(), rtol=2e-2, atol=2e-2)


def test_routed_layer_nodes_on_gpu():
    """ResidualRMSNorm + RoutedFanout + residual-fused RoutedLinear against plain autograd (bf16 tolerances)."""
    need_gpu()
    import libra_b200.functional as LF
    g = _gen(31)
    N, n, C, I = 300, 170, 256, 352
    mk = lambda *s, sc=0.05: (torch.randn(*s, device=dev, generator=g) * sc).bfloat16().requires_grad_(True)
    x = (torch.randn(N, C, device=dev, generator=g)).bfloat16().requires_grad_(True)
    wl, wv = mk(C, sc=1.0), mk(C, sc=1.0)
    flag = torch.zeros(N, dtype=torch.uint8, device=dev)
    flag[n:] = 1
    W, A, B = mk(C, C), mk(C // 4, C), mk(C, C // 4)
    W2, A2, B2 = mk(I, C), mk(I // 4, C), mk(I, I // 4)
    Al, Av = mk(8, C), mk(8, C)
    leaves = [x, wl, wv, W, A, B, W2, A2, B2, Al, Av]

    def run(fused):
        for t in leaves:
            t.grad = None
        if fused:
            hr, n1 = LF.residual_rmsnorm(x, wl, wv, flag, 1e-6)
            y, y2, t = LF.routed_fanout(n1, n, ("lin", "lin", "down"), W, A, B, W2, A2, B2, Al, Av)
            out = LF.routed_linear(y, n, W, A, B, residual=hr)
        else:
            xf = x.float()
            w_rows = torch.where(flag.bool()[:, None], wv.float(), wl.float())
            n1 = (w_rows * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))).bfloat16()
            rl = lambda t_, W_, A_, B_: torch.cat([t_[:n] @ W_.t(), (t_[n:] @ A_.t()) @ B_.t()])
            y, y2 = rl(n1, W, A, B), rl(n1, W2, A2, B2)
            t = torch.cat([n1[:n] @ Al.t(), n1[n:] @ Av.t()])
            out = rl(y, W, A, B) + x
        loss = out.float().pow(2).mean() + y2.float().pow(2).mean() + t.float().pow(2).mean()
        loss.backward()
        return loss.detach(), [t_.grad.clone().float() for t_ in leaves]

    l0, g0 = run(False)
    l1, g1 = run(True)
    assert abs(float(l0) - float(l1)) < 2e-2 * abs(float(l0)) + 1e-3
    for a, b, nm in zip(g1, g0, "x wl wv W A B W2 A2 B2 Al Av".split()):
        assert rel_err(a, b) < 5e-2, (nm, rel_err(a, b))
