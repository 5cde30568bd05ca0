"""N2: the vision tokenizer's decode side (ids -> pixels) on the libra_b200 kernels.

Mirrors, with the reference's parameter names (so a VQModel checkpoint loads as is):
  ImageTokenizer.decode                      libra/models/libra/image_tokenizer.py:97-124
  VQModel.decode_code / decode               libra/models/libra/taming/models/vqgan.py:122-130
  LFQ.indices_to_codes (+ project_out)       taming/modules/quantization/lookup_free_quantization.py:129-158
  Decoder / ResnetBlock / AttnBlock / Upsample / Normalize   taming/modules/diffusionmodules/model.py:29-60, 85-230, 474-588

Inference only (the tokenizer is frozen and runs under no_grad in the reference, image_tokenizer.py:37-42, 97).
Data layout and the "3x3 convolution = one nine-segment GEMM over row shifts" formulation: csrc/vqdec.cu.  Every dense
product (1x1 / 3x3 convolutions, q.k^T, p.v) runs on lb_gemm_grouped; GroupNorm(+swish), nearest upsample, padding and the
row softmax are the kernels of csrc/vqdec.cu.  There is no eager fallback: without the CUDA library the calls raise.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from .. import _lib, ops

BF16 = torch.bfloat16


def Normalize(c: int) -> nn.GroupNorm:          # model.py:34-35
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):                   # parameter container (model.py:85-138); compute is in VQDecoder
    def __init__(self, in_channels: int, out_channels: int, conv_shortcut: bool = False):
        super().__init__()
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = Normalize(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(nn.Module):                     # model.py:141-230
    def __init__(self, in_channels: int, num_attn_head: int = 1):
        super().__init__()
        self.in_channels, self.num_attn_head = in_channels, num_attn_head
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)


class Upsample(nn.Module):                      # model.py:44-60
    def __init__(self, in_channels: int, with_conv: bool, scale_factor: float = 2.0):
        super().__init__()
        self.with_conv, self.scale_factor = with_conv, scale_factor
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)


class Decoder(nn.Module):
    """taming Decoder (model.py:474-588): same constructor keywords, same module tree / state-dict keys."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0, resamp_with_conv=True,
                 in_channels=3, resolution, z_channels, give_pre_end=False, initial_resolution=None, num_attn_head=1,
                 norm_first=False, **ignorekwargs):
        super().__init__()
        if give_pre_end:
            raise NotImplementedError("Decoder(give_pre_end=True) is not part of the decode path")
        self.ch, self.out_ch, self.num_resolutions, self.num_res_blocks = ch, out_ch, len(ch_mult), num_res_blocks
        self.resolution, self.norm_first, self.num_attn_head = resolution, norm_first, num_attn_head
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = initial_resolution if initial_resolution is not None else resolution // 2 ** (self.num_resolutions - 1)
        self.initial_resolution = curr_res
        self.z_shape = (1, z_channels, curr_res, curr_res)
        if norm_first:
            self.first_norm = Normalize(z_channels)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in, num_attn_head)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in, num_attn_head))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level > 1:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            elif i_level == 1:
                up.upsample = Upsample(block_in, resamp_with_conv, scale_factor=resolution / curr_res)
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)


def nearest_source_index(in_size: int, scale_factor: float) -> torch.Tensor:
    """Source index of every output index of F.interpolate(mode="nearest", scale_factor=s) (recompute_scale_factor unset):
    out = floor(in * s), src = min(floor(dst * (1 / s)), in - 1) evaluated in fp32 like ATen's
    nearest_neighbor_compute_source_index."""
    out = int(math.floor(float(in_size) * scale_factor))
    scale = torch.tensor(1.0 / scale_factor, dtype=torch.float32)
    src = torch.floor(torch.arange(out, dtype=torch.float32) * scale).to(torch.int64).clamp_(max=in_size - 1)
    return src.to(torch.int32)


class _Act:
    """An NHWC bf16 activation: padded-row layout with W+3 guard rows on both sides (csrc/vqdec.cu) or compact."""

    def __init__(self, B, H, W, C, device, padded=True):
        self.B, self.H, self.W, self.C, self.padded = B, H, W, C, padded
        if padded:
            self.guard = W + 3
            self.rows = B * (H + 2) * (W + 2)
            self.buf = torch.zeros(self.rows + 2 * self.guard, C, dtype=BF16, device=device)
            self.body = self.buf[self.guard:self.guard + self.rows]
        else:
            self.guard, self.rows = 0, B * H * W
            self.buf = self.body = torch.empty(self.rows, C, dtype=BF16, device=device)

    def shifted(self, shift: int) -> torch.Tensor:
        return self.buf[self.guard + shift:self.guard + shift + self.rows]


def _p(t):
    return None if t is None else t.data_ptr()


class VQDecoder(nn.Module):
    """VQModel's decode side: quantize.project_out (optional) -> post_quant_conv -> Decoder, driven from token ids.

    `ddconfig` = the Decoder keywords of vision_tokenizer_config.yaml (params.ddconfig); embed_dim / codebook_size /
    num_codebook as in VisionTokenizer.  State-dict keys: decoder.*, post_quant_conv.*, quantize.project_out.* (only when
    embed_dim != num_codebook * log2(codebook_size), lookup_free_quantization.py:84-88)."""

    def __init__(self, ddconfig: Dict, embed_dim: int = 18, codebook_size: int = 512, num_codebook: int = 2, token_offset: int = 32000):
        super().__init__()
        self.decoder = Decoder(**ddconfig)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.bits = int(math.log2(codebook_size))
        self.codebook_size, self.num_codebook, self.offset = codebook_size, num_codebook, token_offset
        self.boi_token_id = token_offset + codebook_size
        code_dims = self.bits * num_codebook
        self.quantize = nn.Module()
        if code_dims != embed_dim:
            self.quantize.project_out = nn.Linear(code_dims, embed_dim)
        self._packed = None
        self.requires_grad_(False)
        self.eval()

    def train(self, mode=True):
        return super().train(False)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------ weights, repacked once
    @staticmethod
    def _pad8(n: int) -> int:
        return (n + 7) // 8 * 8

    def _pack(self):
        """conv weight [Co,Ci,3,3] -> [9, Co8, Ci8] (tap-major, `nn.Linear` layout per tap); 1x1 / Linear -> [Co8, Ci8]; bias
        -> [Co8]; zero padding to multiples of 8 (a GEMM row pitch is a multiple of 16 bytes)."""
        if self._packed is not None:
            return self._packed
        dev = self.post_quant_conv.weight.device
        P: Dict[str, torch.Tensor] = {}
        for name, m in self.named_modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                w = m.weight.detach()
                co, ci = w.shape[0], w.shape[1]
                co8, ci8 = self._pad8(co), self._pad8(ci)
                if w.dim() == 4 and w.shape[-1] == 3:
                    t = torch.zeros(9, co8, ci8, dtype=BF16, device=dev)
                    t[:, :co, :ci] = w.permute(2, 3, 0, 1).reshape(9, co, ci).to(BF16)
                else:
                    t = torch.zeros(co8, ci8, dtype=BF16, device=dev)
                    t[:co, :ci] = w.reshape(co, ci).to(BF16)
                b = torch.zeros(co8, dtype=BF16, device=dev)
                b[:co] = m.bias.detach().to(BF16)
                P[name + ".w"], P[name + ".b"] = t, b
            elif isinstance(m, nn.GroupNorm):
                P[name + ".w"], P[name + ".b"] = m.weight.detach().to(BF16).contiguous(), m.bias.detach().to(BF16).contiguous()
        self._packed = P
        return P

    # ------------------------------------------------------------------ ops
    def _gn(self, name: str, x: _Act, swish: bool, out_padded: bool = True) -> _Act:
        P = self._packed
        y = _Act(x.B, x.H, x.W, x.C, x.buf.device, padded=out_padded)
        nchunk = _lib.load().lb_vq_groupnorm_chunks(x.H, x.W)
        ws = torch.empty(x.B * nchunk * 2 * x.C, dtype=torch.float32, device=x.buf.device)
        _lib.call("lb_vq_groupnorm", _p(x.body), _p(P[name + ".w"]), _p(P[name + ".b"]), _p(y.body), _p(ws), x.B, x.H, x.W, x.C, 32,
                  1e-6, int(swish), int(x.padded), int(out_padded), ops._st())
        return y

    def _conv3(self, name: str, x: _Act, d: Optional[_Act] = None) -> _Act:
        """3x3 / stride 1 / pad 1 convolution of a padded activation (zero borders) -> padded activation (garbage borders)."""
        assert x.padded
        w9, b = self._packed[name + ".w"], self._packed[name + ".b"]
        y = _Act(x.B, x.H, x.W, w9.shape[1], x.buf.device)
        es = []
        for t in range(9):
            dy, dx = divmod(t, 3)
            a = x.shifted((dy - 1) * (x.W + 2) + (dx - 1))
            es.append(ops.gp(a, w9[t], y.body, bias=b if t == 0 else None, d=(d.body if (d is not None and t == 0) else None),
                             acc_prev=t > 0))
        ops.gemm_grouped(es)
        return y

    def _conv1(self, name: str, x: _Act, d: Optional[_Act] = None) -> _Act:
        w, b = self._packed[name + ".w"], self._packed[name + ".b"]
        y = _Act(x.B, x.H, x.W, w.shape[0], x.buf.device, padded=x.padded)
        ops.gemm_grouped([ops.gp(x.body, w, y.body, bias=b, d=None if d is None else d.body)])
        return y

    def _resnet(self, name: str, m: ResnetBlock, x: _Act) -> _Act:
        h = self._conv3(name + ".conv1", self._gn(name + ".norm1", x, True))
        h = self._gn(name + ".norm2", h, True)
        if m.in_channels != m.out_channels:
            if m.use_conv_shortcut:                                       # Decoder never passes conv_shortcut=True (model.py:527-531)
                raise NotImplementedError("ResnetBlock(conv_shortcut=True) is not used by the taming Decoder")
            x = self._conv1(name + ".nin_shortcut", x)
        return self._conv3(name + ".conv2", h, d=x)                      # x + h: the residual rides in the GEMM epilogue

    def _attn(self, name: str, m: AttnBlock, x: _Act) -> _Act:
        if not x.padded:
            raise NotImplementedError("attention on a compact activation")
        B, H, W, C = x.B, x.H, x.W, x.C
        heads, N = m.num_attn_head, x.H * x.W
        dh = C // heads
        n = self._gn(name + ".norm", x, False, out_padded=False)
        P = self._packed
        q, k, v = (torch.empty(B * N, C, dtype=BF16, device=n.buf.device) for _ in range(3))
        ops.gemm_grouped([ops.gp(n.body, P[f"{name}.{t}.w"], o, bias=P[f"{name}.{t}.b"]) for t, o in (("q", q), ("k", k), ("v", v))])
        Np = self._pad8(N)
        s = torch.empty(B * heads, N, Np, dtype=BF16, device=q.device)
        o = torch.empty(B * N, C, dtype=BF16, device=q.device)
        sl = lambda t, b, h: t[b * N:(b + 1) * N, h * dh:(h + 1) * dh]
        pairs = [(b, h) for b in range(B) for h in range(heads)]
        for i in range(0, len(pairs), 16):                                # w_[i, j] = q_i . k_j  (model.py:203-206)
            ops.gemm_grouped([ops.gp(sl(q, b, h), sl(k, b, h), s[b * heads + h][:, :N]) for b, h in pairs[i:i + 16]])
        _lib.call("lb_softmax_rows", _p(s), B * heads * N, N, Np, float(int(dh) ** -0.5), ops._st())
        for i in range(0, len(pairs), 16):                                # h_[i] = sum_j w_[i, j] v_j  (:212-215)
            ops.gemm_grouped([ops.gp(s[b * heads + h][:, :N], sl(v, b, h), sl(o, b, h), tb=True) for b, h in pairs[i:i + 16]])
        pr = torch.empty(B * N, C, dtype=BF16, device=q.device)
        ops.gemm_grouped([ops.gp(o, P[name + ".proj_out.w"], pr, bias=P[name + ".proj_out.b"])])
        y = _Act(B, H, W, C, q.device)
        _lib.call("lb_vq_pad", _p(pr), C, _p(x.body), _p(y.body), B, H, W, C, ops._st())      # x + proj_out(h_), borders zeroed
        return y

    def _upsample(self, name: str, m: Upsample, x: _Act) -> _Act:
        sy = nearest_source_index(x.H, m.scale_factor).to(x.buf.device)
        sx = nearest_source_index(x.W, m.scale_factor).to(x.buf.device)
        y = _Act(x.B, sy.numel(), sx.numel(), x.C, x.buf.device)
        _lib.call("lb_vq_upsample_nearest", _p(x.body), _p(y.body), _p(sy), _p(sx), x.B, x.H, x.W, x.C, y.H, y.W, int(x.padded), ops._st())
        return self._conv3(name + ".conv", y) if m.with_conv else y

    # ------------------------------------------------------------------ the reference entry points
    @torch.no_grad()
    def decode_ids(self, ids: torch.Tensor) -> torch.Tensor:
        """ImageTokenizer.decode: ids [Q,B,N] or [B,N]-less [Q,N] (with or without <img>/</img>) -> pixels [B,out_ch,R,R] bf16."""
        _lib.require_device()
        if ids.dim() == 2:
            ids = ids[None]
        if ids.dim() != 3:
            raise NotImplementedError
        if bool((ids == self.boi_token_id).any()):
            ids = ids[:, :, 1:-1]
        Q, B, N = ids.shape
        side = int(math.isqrt(N))
        if side * side != N:
            raise ValueError("Input images are invalid. Currently, the image decoder only support square images.")
        P = self._pack()
        dev = ids.device
        kd = self._pad8(Q * self.bits)
        codes = torch.empty(B * N, kd, dtype=BF16, device=dev)
        _lib.call("lb_vq_codes", _p(ids.contiguous()), int(self.offset), Q, B * N, self.bits, _p(codes), kd, ops._st())
        z = codes
        if hasattr(self.quantize, "project_out"):
            w, b = P["quantize.project_out.w"], P["quantize.project_out.b"]
            z2 = torch.empty(B * N, w.shape[0], dtype=BF16, device=dev)
            ops.gemm_grouped([ops.gp(z, w, z2, bias=b)])
            z = z2
        w, b = P["post_quant_conv.w"], P["post_quant_conv.b"]
        zc = _Act(B, side, side, w.shape[0], dev, padded=False)
        ops.gemm_grouped([ops.gp(z, w, zc.body, bias=b)])
        return self._decoder_forward(zc)

    decode = decode_ids

    def _decoder_forward(self, z: _Act) -> torch.Tensor:
        d = self.decoder
        if d.norm_first:
            z = self._gn("decoder.first_norm", z, False)
        else:
            zp = _Act(z.B, z.H, z.W, z.C, z.buf.device)
            _lib.call("lb_vq_pad", _p(z.body), z.C, None, _p(zp.body), z.B, z.H, z.W, z.C, ops._st())
            z = zp
        h = self._conv3("decoder.conv_in", z)
        h = self._resnet("decoder.mid.block_1", d.mid.block_1, h)
        h = self._attn("decoder.mid.attn_1", d.mid.attn_1, h)
        h = self._resnet("decoder.mid.block_2", d.mid.block_2, h)
        for lvl in reversed(range(d.num_resolutions)):
            up = d.up[lvl]
            for i in range(d.num_res_blocks + 1):
                h = self._resnet(f"decoder.up.{lvl}.block.{i}", up.block[i], h)
                if len(up.attn) > 0:
                    h = self._attn(f"decoder.up.{lvl}.attn.{i}", up.attn[i], h)
            if lvl != 0:
                h = self._upsample(f"decoder.up.{lvl}.upsample", up.upsample, h)
        h = self._conv3("decoder.conv_out", self._gn("decoder.norm_out", h, True))
        out = torch.empty(h.B, d.out_ch, h.H, h.W, dtype=BF16, device=h.buf.device)
        _lib.call("lb_vq_to_nchw", _p(h.body), _p(out), h.B, h.H, h.W, h.C, d.out_ch, ops._st())
        return out
