"""N2 on the GPU: the vision tokenizer's decode path (ids -> pixels) against the fixture produced by the reference's own
taming Decoder + LFQ.indices_to_codes (tests/golden/vq_decode_tiny.pt, oracle/make_golden.py:golden_vq_decode), against the
oracle in bf16, and kernel by kernel against PyTorch."""
import math

import pytest
import torch
import torch.nn.functional as F

from gpu_util import need_gpu, assert_close, rel_err
from oracle import libra_oracle as O

pytestmark = pytest.mark.gpu
dev = "cuda"
BF16 = torch.bfloat16


def _padded(x_nchw):
    """NCHW fp32/bf16 -> padded-row NHWC bf16 buffer with guards; returns (_Act-like body tensor, B, H, W, C)."""
    from libra_b200.models.vq_decoder import _Act
    B, C, H, W = x_nchw.shape
    a = _Act(B, H, W, C, x_nchw.device)
    a.body.view(B, H + 2, W + 2, C)[:, 1:-1, 1:-1] = x_nchw.permute(0, 2, 3, 1).to(BF16)
    return a


def _interior(a):
    return a.body.view(a.B, a.H + 2, a.W + 2, a.C)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float()


def test_vq_codes_bit_exact():
    need_gpu()
    from libra_b200 import _lib, ops
    g = torch.Generator().manual_seed(0)
    Q, B, N, bits = 2, 3, 36, 9
    ids = (torch.randint(0, 512, (Q, B, N), generator=g) + 32000).to(dev)
    codes = torch.full((B * N, 24), 7.0, dtype=BF16, device=dev)
    _lib.call("lb_vq_codes", ids.data_ptr(), 32000, Q, B * N, bits, codes.data_ptr(), 24, ops._st())
    want = O.lfq_indices_to_codes((ids - 32000).reshape(Q, B, 6, 6).permute(1, 2, 3, 0), bits)      # [B, 18, 6, 6]
    want = want.permute(0, 2, 3, 1).reshape(B * N, Q * bits)
    assert torch.equal(codes[:, :18].float(), want.float())
    assert bool((codes[:, 18:] == 0).all())


@pytest.mark.parametrize("C,H,W,swish,inp,outp", [(64, 6, 6, True, True, True), (32, 12, 12, False, True, False),
                                                  (128, 48, 40, True, False, True), (512, 24, 24, True, True, True)])
def test_groupnorm_swish(C, H, W, swish, inp, outp):
    need_gpu()
    from libra_b200 import _lib, ops
    from libra_b200.models.vq_decoder import _Act
    g = torch.Generator(device=dev).manual_seed(C + H)
    B = 2
    x = (torch.randn(B, C, H, W, device=dev, generator=g) * 1.5 + 0.3).bfloat16()
    gamma = (1 + 0.2 * torch.randn(C, device=dev, generator=g)).bfloat16()
    beta = (0.1 * torch.randn(C, device=dev, generator=g)).bfloat16()
    if inp:
        xa = _padded(x)
        xa.body.view(B, H + 2, W + 2, C)[:, 0] = 9.0           # garbage borders must not reach the statistics
    else:
        xa = _Act(B, H, W, C, dev, padded=False)
        xa.body.copy_(x.permute(0, 2, 3, 1).reshape(-1, C))
    ya = _Act(B, H, W, C, dev, padded=outp)
    if outp:
        ya.body.fill_(5.0)
    nchunk = _lib.load().lb_vq_groupnorm_chunks(H, W)
    ws = torch.empty(B * nchunk * 2 * C, dtype=torch.float32, device=dev)
    _lib.call("lb_vq_groupnorm", xa.body.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ya.body.data_ptr(), ws.data_ptr(), B, H, W, C,
              32, 1e-6, int(swish), int(inp), int(outp), ops._st())
    want = F.group_norm(x.float(), 32, gamma.float(), beta.float(), eps=1e-6)
    if swish:
        want = want * torch.sigmoid(want)
    got = _interior(ya) if outp else ya.body.view(B, H, W, C).permute(0, 3, 1, 2).float()
    assert_close(got, want, rtol=1e-2, atol=2e-2)
    if outp:
        full = ya.body.view(B, H + 2, W + 2, C)
        assert bool((full[:, 0] == 0).all() and (full[:, -1] == 0).all() and (full[:, :, 0] == 0).all() and (full[:, :, -1] == 0).all())


@pytest.mark.parametrize("Cin,Cout,H,W", [(64, 64, 6, 6), (32, 64, 12, 10), (128, 8, 48, 48), (24, 32, 5, 7)])
def test_conv3x3_as_nine_segment_gemm(Cin, Cout, H, W):
    """3x3 / pad 1 convolution = one grouped-GEMM problem of nine K segments over row shifts of the padded layout."""
    need_gpu()
    from libra_b200 import ops
    from libra_b200.models.vq_decoder import _Act
    g = torch.Generator(device=dev).manual_seed(Cin * 7 + Cout)
    B = 2
    x = torch.randn(B, Cin, H, W, device=dev, generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev, generator=g) * 0.1).bfloat16()
    b = torch.randn(Cout, device=dev, generator=g).bfloat16()
    res = torch.randn(B, Cout, H, W, device=dev, generator=g).bfloat16()
    xa, ra = _padded(x), _padded(res)
    w9 = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    ya = _Act(B, H, W, Cout, dev)
    es = []
    for t in range(9):
        dy, dx = divmod(t, 3)
        es.append(ops.gp(xa.shifted((dy - 1) * (W + 2) + (dx - 1)), w9[t], ya.body, bias=b if t == 0 else None,
                         d=ra.body if t == 0 else None, acc_prev=t > 0))
    ops.gemm_grouped(es)
    want = F.conv2d(x.float(), w.float(), b.float(), padding=1).bfloat16().float() + res.float()
    assert_close(_interior(ya), want, rtol=1e-2, atol=3e-2)


@pytest.mark.parametrize("H,scale", [(6, 2.0), (12, 4.0), (24, 336 / 96), (7, 1.5)])
def test_upsample_nearest_matches_interpolate(H, scale):
    need_gpu()
    from libra_b200 import _lib, ops
    from libra_b200.models.vq_decoder import _Act, nearest_source_index
    g = torch.Generator(device=dev).manual_seed(H)
    B, C, W = 2, 32, H + 1
    x = torch.randn(B, C, H, W, device=dev, generator=g).bfloat16()
    want = F.interpolate(x.float(), scale_factor=scale, mode="nearest")
    sy, sx = nearest_source_index(H, scale).to(dev), nearest_source_index(W, scale).to(dev)
    assert (sy.numel(), sx.numel()) == tuple(want.shape[2:])
    xa = _padded(x)
    ya = _Act(B, sy.numel(), sx.numel(), C, dev)
    ya.body.fill_(3.0)
    _lib.call("lb_vq_upsample_nearest", xa.body.data_ptr(), ya.body.data_ptr(), sy.data_ptr(), sx.data_ptr(), B, H, W, C, ya.H, ya.W, 1,
              ops._st())
    assert torch.equal(_interior(ya), want)
    assert bool((ya.body.view(B, ya.H + 2, ya.W + 2, C)[:, 0] == 0).all())


def test_softmax_rows():
    need_gpu()
    from libra_b200 import _lib, ops
    g = torch.Generator(device=dev).manual_seed(1)
    rows, cols, ld = 77, 36, 40
    x = (torch.randn(rows, ld, device=dev, generator=g) * 6).bfloat16()
    ref = x.clone()
    _lib.call("lb_softmax_rows", x.data_ptr(), rows, cols, ld, 0.125, ops._st())
    want = torch.softmax((ref[:, :cols] * 0.125).float(), dim=1)          # bf16 * scale rounds to bf16 first, like eager
    assert_close(x[:, :cols], want, rtol=1e-2, atol=1e-3)
    assert torch.equal(x[:, cols:], ref[:, cols:])


def _decoder(g, dtype=BF16):
    from libra_b200.models.vq_decoder import VQDecoder
    sd = g["state_dict"]
    embed_dim = sd["post_quant_conv.weight"].shape[1]
    m = VQDecoder(dict(g["config"]), embed_dim=embed_dim, codebook_size=g["codebook_size"], num_codebook=g["ids"].shape[0],
                  token_offset=g["token_offset"])
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.to(dtype).to(dev)


def test_vq_decode_matches_reference_golden(golden):
    """ids -> pixels against the pixels the reference's taming Decoder produced (fp32) and the oracle in bf16."""
    need_gpu()
    g = golden("vq_decode_tiny")
    m = _decoder(g)
    ids = g["ids"].to(dev)
    px = m.decode_ids(ids)
    assert px.shape == g["pixels"].shape and px.dtype == BF16
    want = g["pixels"].to(dev)
    sd16 = {k: v.to(dev).to(BF16) for k, v in g["state_dict"].items()}
    dims = O.VQDecoderDims(**{k: v for k, v in g["config"].items() if k in O.VQDecoderDims.__dataclass_fields__})
    orc16 = O.vq_decode(sd16, dims, ids, g["token_offset"], g["codebook_size"], boi_token_id=g["boi_token_id"])
    e_ours, e_orc = rel_err(px, want), rel_err(orc16, want)
    print(f"vq_decode: rel err ours {e_ours:.4g}, bf16 oracle {e_orc:.4g}")
    assert torch.isfinite(px.float()).all()
    assert e_ours <= 1.5 * e_orc + 2e-3, (e_ours, e_orc)
    # ids without <img>/</img> decode to the same pixels (image_tokenizer.py:110-111)
    px2 = m.decode_ids(ids[:, :, 1:-1].contiguous())
    assert torch.equal(px, px2)


def test_vq_decode_wider_config_vs_oracle():
    """A decoder closer to the real one: 128-wide base, multi-head attention at two resolutions, norm_first, three upsamples to
    a non-power-of-two scale; compared with the fp32 oracle (the bf16 oracle as the yardstick)."""
    need_gpu()
    from libra_b200.models.vq_decoder import VQDecoder
    cfg = dict(ch=64, out_ch=3, ch_mult=(1, 2, 4), num_res_blocks=1, attn_resolutions=(8, 16), in_channels=3, resolution=56,
               z_channels=64, initial_resolution=8, num_attn_head=4, norm_first=True)
    torch.manual_seed(3)
    m = VQDecoder(cfg, embed_dim=18, codebook_size=512, num_codebook=2, token_offset=32000)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() > 1:
                p.normal_(0, 1.0 / math.sqrt(p[0].numel()))
            elif "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn_like(p))
            else:
                p.normal_(0, 0.05)
    sd32 = {k: v.detach().clone().to(dev) for k, v in m.state_dict().items()}
    m = m.to(BF16).to(dev)
    g = torch.Generator().manual_seed(5)
    ids = (torch.randint(0, 512, (2, 2, 64), generator=g) + 32000).to(dev)
    px = m.decode_ids(ids)
    dims = O.VQDecoderDims(**{k: v for k, v in cfg.items() if k in O.VQDecoderDims.__dataclass_fields__})
    want = O.vq_decode(sd32, dims, ids, 32000, 512)
    orc16 = O.vq_decode({k: v.to(BF16) for k, v in sd32.items()}, dims, ids, 32000, 512)
    assert px.shape == want.shape == (2, 3, 56, 56)
    e_ours, e_orc = rel_err(px, want), rel_err(orc16, want)
    print(f"vq_decode wide: rel err ours {e_ours:.4g}, bf16 oracle {e_orc:.4g}")
    assert e_ours <= 1.5 * e_orc + 2e-3, (e_ours, e_orc)
