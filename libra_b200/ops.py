"""Raw (non-autograd) Python entry points over the C ABI.  Every function enqueues CUDA work on
torch's current stream and returns torch tensors that own the output memory.

These are the leaves; `libra_b200.functional` wraps them in torch.autograd.Function objects and
`libra_b200.models` composes those into the reference's module interface.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

from . import _lib

BF16 = torch.bfloat16
DT_BF16, DT_F32 = 0, 1


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# Optional device-side timing of selected kernels (bench.py roofline): name -> list of (start_event, stop_event)
# recorded on torch's current stream, i.e. the stream the kernel is launched on.
TIMED = None


STREAM_HEAD_GROUP = 8          # heads per CTA-order group of the persistent attention forward (see schedule.stream_plan)


_SM_COUNT = None


def sm_count() -> int:
    global _SM_COUNT
    if _SM_COUNT is None:
        _lib.require_device()
        _SM_COUNT = int(_lib.load().lb_sm_count())
    return _SM_COUNT


class pdl:
    """Context manager: programmatic dependent launch for the one-token decode chain (lb_set_pdl; see include/libra_b200.h).
    LB_PDL=0 in the environment keeps the launches serial (A/B runs)."""

    def __init__(self, on: bool = True):
        import os
        self.on = bool(on) and os.environ.get("LB_PDL", "1") != "0"

    def __enter__(self):
        self.prev = int(_lib.load().lb_set_pdl(1 if self.on else 0))
        return self

    def __exit__(self, *exc):
        _lib.load().lb_set_pdl(self.prev)
        return False


def enable_timing(names=("lb_attn_fwd", "lb_attn_fwd_stream", "lb_attn_bwd_dq", "lb_attn_bwd_dq_stream", "lb_attn_bwd_dkv", "lb_attn_bwd_dkv_stream")):
    global TIMED
    TIMED = {n: [] for n in names}


def disable_timing():
    global TIMED
    t, TIMED = TIMED, None
    return t


def _timed_call(name, *args):
    if TIMED is not None and name in TIMED:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        _lib.call(name, *args)
        e.record()
        TIMED[name].append((s, e))
    else:
        _lib.call(name, *args)


def _chk(t: torch.Tensor, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise _lib.LibraB200Error(f"{name} must be a CUDA tensor (libra_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


# ------------------------------------------------------------------ norms
def rmsnorm_fwd(x, w_lang, w_vis=None, flag=None, eps=1e-6):
    _chk(x, BF16, "x"); _chk(w_lang, BF16, "w_lang")
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    y = torch.empty_like(x)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    _lib.call("lb_rmsnorm_fwd", _p(x), _p(w_lang), _p(w_vis), _p(flag), _p(y), _p(rstd), rows, cols, float(eps), _st())
    return y, rstd


def rmsnorm_bwd(dy, x, w_lang, w_vis, flag, rstd, residual_grad=None, need_dw=True):
    _chk(dy, BF16, "dy"); _chk(x, BF16, "x")
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    dx = torch.empty_like(x)
    ws = torch.empty(_lib.load().lb_rmsnorm_bwd_workspace(rows, cols), dtype=torch.uint8, device=x.device)
    dwl = torch.zeros(cols, dtype=torch.float32, device=x.device) if need_dw else None
    dwv = torch.zeros(cols, dtype=torch.float32, device=x.device) if (need_dw and w_vis is not None) else None
    _lib.call("lb_rmsnorm_bwd", _p(dy), _p(x), _p(w_lang), _p(w_vis), _p(flag), _p(rstd), _p(residual_grad), _p(dx),
              _p(dwl), _p(dwv), _p(ws), rows, cols, _st())
    return dx, dwl, dwv


def layernorm_fwd(x, w, b, eps=1e-5):
    _chk(x, BF16, "x")
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    y = torch.empty_like(x)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    _lib.call("lb_layernorm_fwd", _p(x), _p(w), _p(b), _p(y), _p(mean), _p(rstd), rows, cols, float(eps), _st())
    return y, mean, rstd


def layernorm_bwd(dy, x, w, mean, rstd):
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    dx = torch.empty_like(x)
    ws = torch.empty(_lib.load().lb_layernorm_bwd_workspace(rows, cols), dtype=torch.uint8, device=x.device)
    dw = torch.zeros(cols, dtype=torch.float32, device=x.device)
    db = torch.zeros(cols, dtype=torch.float32, device=x.device)
    _lib.call("lb_layernorm_bwd", _p(dy), _p(x), _p(w), _p(mean), _p(rstd), _p(dx), _p(dw), _p(db), _p(ws), rows, cols, _st())
    return dx, dw, db


# ------------------------------------------------------------ elementwise
def swiglu_fwd(gate, up, out=None):
    """gate/up/out: [rows, cols] views (last dim contiguous, row pitch arbitrary multiple of 8)."""
    rows, cols = gate.shape
    if out is None:
        out = torch.empty(rows, cols, dtype=BF16, device=gate.device)
    if rows > 0:
        _lib.call("lb_swiglu_fwd", _p(gate), _p(up), _p(out), rows, cols, gate.stride(0), up.stride(0), out.stride(0), _st())
    return out


def swiglu_bwd(dout, gate, up, dgate=None, dup=None):
    rows, cols = gate.shape
    if dgate is None:
        dgate = torch.empty(rows, cols, dtype=BF16, device=gate.device)
    if dup is None:
        dup = torch.empty(rows, cols, dtype=BF16, device=gate.device)
    _lib.call("lb_swiglu_bwd", _p(dout), _p(gate), _p(up), _p(dgate), _p(dup), rows, cols, dout.stride(0),
              gate.stride(0), up.stride(0), dgate.stride(0), dup.stride(0), _st())
    return dgate, dup


def bias_quick_gelu_fwd(x, bias=None):
    _chk(x, BF16, "x")
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    y = torch.empty_like(x)
    _lib.call("lb_bias_quick_gelu_fwd", _p(x), _p(bias), _p(y), rows, cols, _st())
    return y


def bias_quick_gelu_bwd(dy, x, bias=None):
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    dx = torch.empty_like(x)
    _lib.call("lb_bias_quick_gelu_bwd", _p(dy), _p(x), _p(bias), _p(dx), rows, cols, _st())
    return dx


def gather_rows(src, index):
    """dst[r] = src[index[r]]; src [*, cols] bf16, index int32."""
    _chk(src, BF16, "src"); _chk(index, torch.int32, "index")
    cols = src.shape[-1]
    dst = torch.empty(index.numel(), cols, dtype=BF16, device=src.device)
    _lib.call("lb_gather_rows", _p(src), _p(index), _p(dst), index.numel(), cols, _st())
    return dst


def embed_lang(ids, table):
    _chk(ids, torch.int64, "ids"); _chk(table, BF16, "table")
    out = torch.empty(ids.numel(), table.shape[1], dtype=BF16, device=table.device)
    _lib.call("lb_embed_lang_fwd", _p(ids), _p(table), _p(out), ids.numel(), table.shape[1], _st())
    return out


def embed_vision_cat(ids0, ids1, table0, table1, signal, signal_row, signal_cols):
    half = table0.shape[1]
    rows = ids0.numel()
    out = torch.empty(rows, 2 * half + signal_cols, dtype=BF16, device=table0.device)
    _lib.call("lb_embed_vision_cat_fwd", _p(ids0), _p(ids1), _p(table0), _p(table1), _p(signal), _p(signal_row), _p(out),
              rows, half, signal_cols, _st())
    return out


def embed_bwd(ids, dy, col0, cols, dtable):
    """dtable[ids[r]] += dy[r, col0:col0+cols]; dtable fp32."""
    _lib.call("lb_embed_bwd", _p(ids), _p(dy), dy.stride(0), col0, _p(dtable), ids.numel(), cols, _st())
    return dtable


def lfq_pack(h, n_img, tokens, num_codebooks, bits, offset, boi, eoi):
    dt = DT_BF16 if h.dtype == BF16 else DT_F32
    if h.dtype not in (BF16, torch.float32):
        raise TypeError("lfq_pack: bf16 or fp32")
    _chk(h, None, "h")
    ids = torch.empty(num_codebooks, n_img, tokens + 2, dtype=torch.int64, device=h.device)
    _lib.call("lb_lfq_pack", _p(h), dt, n_img, tokens, num_codebooks, bits, offset, boi, eoi, _p(ids), _st())
    return ids


def lfq_unpack(idx, num_codebooks, bits, dtype=BF16):
    _chk(idx, torch.int64, "idx")
    n = idx.numel() // num_codebooks
    codes = torch.empty(*idx.shape[:-1], num_codebooks * bits, dtype=dtype, device=idx.device)
    _lib.call("lb_lfq_unpack", _p(idx), n, num_codebooks, bits, _p(codes), DT_BF16 if dtype == BF16 else DT_F32, _st())
    return codes


def cross_entropy_fwd_bwd(logits, labels, vocab, grad_scale):
    """In place: logits (bf16 [rows, ld]) become grad_scale * d(sum loss)/dlogits.  Returns per-row loss (fp32)."""
    rows = logits.shape[0]
    loss = torch.empty(rows, dtype=torch.float32, device=logits.device)
    _lib.call("lb_cross_entropy_fwd_bwd", _p(logits), logits.stride(0), _p(labels), _p(loss), rows, vocab,
              float(grad_scale), _st())
    return loss


# ---------------------------------------------------------------- GEMM
def gemm(a, b, trans_a=False, trans_b=False, out=None, out_dtype=BF16, bias=None, act=0, accumulate=False):
    """C = op(A) op(B).  a: [M,K] (or [K,M] if trans_a); b: [N,K] (or [K,N] if trans_b) -- nn.Linear layout by default."""
    M = a.shape[1] if trans_a else a.shape[0]
    K = a.shape[0] if trans_a else a.shape[1]
    N = b.shape[1] if trans_b else b.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    _lib.call("lb_gemm_bf16", _p(a), _p(b), _p(out), _p(bias), M, N, K, a.stride(0), b.stride(0), out.stride(0),
              int(trans_a), int(trans_b), DT_BF16 if out.dtype == BF16 else DT_F32, int(accumulate), int(act), _st())
    return out


class GemmProblem(ctypes.Structure):
    """lb_gemm_problem (include/libra_b200.h)."""
    _fields_ = [("A", ctypes.c_void_p), ("B", ctypes.c_void_p), ("C", ctypes.c_void_p), ("D", ctypes.c_void_p),
                ("bias", ctypes.c_void_p), ("B2", ctypes.c_void_p), ("G", ctypes.c_void_p), ("U", ctypes.c_void_p),
                ("M", ctypes.c_int64), ("N", ctypes.c_int64), ("K", ctypes.c_int64),
                ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64), ("ldc", ctypes.c_int64), ("ldd", ctypes.c_int64),
                ("trans_a", ctypes.c_int32), ("trans_b", ctypes.c_int32), ("epilogue", ctypes.c_int32),
                ("wait_on", ctypes.c_int32), ("alpha", ctypes.c_void_p), ("flags", ctypes.c_int64)]


EPI_NONE, EPI_QGELU, EPI_SWIGLU = 0, 1, 2


def _ld(t: torch.Tensor, name: str) -> int:
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1) or t.dtype != BF16 or not t.is_cuda:
        raise ValueError(f"gemm operand {name}: need a 2-D bf16 CUDA tensor with unit inner stride, got "
                         f"{tuple(t.shape)} strides {t.stride()} {t.dtype} {t.device}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gp(a, b, c, *, ta=False, tb=False, d=None, bias=None, epi=EPI_NONE, b2=None, g=None, u=None, wait_on=-1, alpha=None,
       acc_prev=False):
    """One entry of a grouped launch:  c[M,N] = epi(alpha * op(a) . op(b) [+ bias]) [+ d].
    a: [M,K] (ta: [K,M]);  b: [N,K], the nn.Linear weight layout (tb: [K,N]);  c, d, g, u: [M,N] row-major views.
    acc_prev: this product is summed into the previous entry's accumulator instead (c = that entry's c).
    wait_on: index (in the launch's list) of the entry that produces `a`."""
    M = a.shape[1] if ta else a.shape[0]
    K = a.shape[0] if ta else a.shape[1]
    N = b.shape[1] if tb else b.shape[0]
    if (b.shape[0] if tb else b.shape[1]) != K or tuple(c.shape) != (M, N):
        raise ValueError(f"gemm shape mismatch: a {tuple(a.shape)} ta={ta}, b {tuple(b.shape)} tb={tb}, c {tuple(c.shape)}")
    q = GemmProblem()
    q.A, q.B, q.C = a.data_ptr(), b.data_ptr(), c.data_ptr()
    q.M, q.N, q.K = M, N, K
    q.lda, q.ldb, q.ldc = _ld(a, "a"), _ld(b, "b"), _ld(c, "c")
    q.trans_a, q.trans_b, q.epilogue, q.wait_on = int(ta), int(tb), int(epi), int(wait_on)
    if d is not None:
        if tuple(d.shape) != (M, N):
            raise ValueError("gemm addend shape")
        q.D, q.ldd = d.data_ptr(), _ld(d, "d")
    if bias is not None:
        q.bias = bias.data_ptr()
    if alpha is not None:
        if alpha.dtype != torch.float32 or alpha.numel() != 1 or not alpha.is_cuda:
            raise ValueError("gemm alpha: a CUDA fp32 scalar tensor")
        q.alpha = alpha.data_ptr()
    q.flags = 1 if acc_prev else 0
    if b2 is not None:
        if tuple(b2.shape) != tuple(b.shape) or _ld(b2, "b2") != q.ldb:
            raise ValueError("gemm b2 must match b")
        q.B2 = b2.data_ptr()
    for nm, t in (("G", g), ("U", u)):
        if t is not None:
            if tuple(t.shape) != (M, N) or _ld(t, nm) != q.ldc:
                raise ValueError(f"gemm extra output {nm} must match c (shape and pitch)")
            setattr(q, nm, t.data_ptr())
    return q


_GG_WS = {}


SKINNY_MAX_M = 32                # rows up to which a plain x W^T product takes the weight-streaming kernel (csrc/gemm_skinny.cu)
SKINNY = os.environ.get("LB_GEMM_SKINNY", "1") != "0"
_SK_WS = {}


def _skinny_ok(q) -> bool:
    return (1 <= q.M <= SKINNY_MAX_M and not q.trans_a and not q.trans_b and q.wait_on < 0 and not q.alpha and not q.G and not q.U
            and not (q.flags & 1) and (q.epilogue == EPI_NONE or (q.epilogue == EPI_SWIGLU and not q.bias)) and q.K % 8 == 0)


def gemm_skinny(problems):
    """Up to 8 plain products with M <= 32 rows per launch on the weight-streaming kernel (csrc/gemm_skinny.cu)."""
    lib = _lib.load()
    dev = torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream().cuda_stream)
    for i in range(0, len(problems), 8):
        chunk = problems[i:i + 8]
        n = len(chunk)
        arr = (GemmProblem * n)(*chunk)
        need = int(lib.lb_gemm_skinny_workspace_bytes(arr, n))
        if need < 0:
            raise _lib.LibraB200Error(f"lb_gemm_skinny_workspace_bytes failed: {_lib.last_error()}")
        ws = _SK_WS.get(key)
        if ws is None or ws.numel() < need:         # zeroed once: the kernel leaves its tile counters at zero
            ws = _SK_WS[key] = torch.zeros(max(need, 16 << 20), dtype=torch.uint8, device=f"cuda:{dev}")
        _timed_call("lb_gemm_skinny", arr, n, _p(ws), ws.numel(), _st())


def gemm_grouped(problems):
    """Run up to 16 problems (see gp()) as ONE persistent tcgen05 launch on the current stream (csrc/gemm_grouped.cu).
    A list made only of plain products with M <= 32 rows (the one-token decode step) goes to gemm_skinny instead."""
    n = len(problems)
    if n == 0:
        return
    if SKINNY and all(_skinny_ok(q) for q in problems):
        return gemm_skinny(problems)
    arr = (GemmProblem * n)(*problems)
    dev = torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream().cuda_stream)
    ws = _GG_WS.get(key)
    if ws is None:                          # per-stream tile counter + chain counters (zeroed by the library per launch)
        ws = _GG_WS[key] = torch.zeros(4096, dtype=torch.int32, device=f"cuda:{dev}")
    _timed_call("lb_gemm_grouped", arr, n, _p(ws), 0 if ws is None else ws.numel() * 4, _st())


def linear(x, w, out=None, *, bias=None, epi=EPI_NONE, d=None, g=None):
    """y = x w^T (+bias, activation, addend) through the grouped kernel (single problem)."""
    if out is None:
        out = torch.empty(x.shape[0], w.shape[0], dtype=BF16, device=x.device)
    gemm_grouped([gp(x, w, out, bias=bias, epi=epi, d=d, g=g)])
    return out


def patch_embed_pack_weight(weight):
    """conv weight [C,3,14,14] (bf16) -> K-packed [C,768]."""
    C = weight.shape[0]
    packed = torch.empty(C, 768, dtype=BF16, device=weight.device)
    _lib.call("lb_patch_embed_pack_weight", _p(weight.contiguous()), _p(packed), C, weight.shape[-1], _st())
    return packed


def patch_embed_fwd(pixels, weight_packed, class_emb, pos_emb, patch=14):
    """pixels [B,3,S,S] bf16 -> embeddings [B, (S/patch)^2+1, C] (class token + position embedding added)."""
    _chk(pixels, BF16, "pixels")
    B, _, S, _ = pixels.shape
    C = weight_packed.shape[0]
    G = S // patch
    emb = torch.empty(B, G * G + 1, C, dtype=BF16, device=pixels.device)
    _lib.call("lb_patch_embed_fwd", _p(pixels), _p(weight_packed), _p(class_emb), _p(pos_emb), _p(emb), B, S, patch, C, _st())
    return emb


def probe_umma(mode, a, b):
    K = a.shape[1]
    d = torch.empty(128, 128, dtype=torch.float32, device=a.device)
    _lib.call("lb_probe_umma", mode, _p(a), _p(b), _p(d), K, _st())
    return d


# ------------------------------------------------------------ attention
def attn_prep_fwd(q, k, kc, v, vc, flag_sorted, sorted_of, pos, cos_t, sin_t, heads, head_dim, kv_out=None, kv_row=None):
    """kc/vc: bridged variants (k + kb, v + vb) or None.  kv_out = (Kfv, Kfl, Vfv, Vfl) destination tensors with row map
    kv_row (decode: the KV cache and the tokens' slots); default: fresh [n, C] tensors, identity rows."""
    n = q.shape[0]
    C = heads * head_dim
    Q = torch.empty(n, C, dtype=BF16, device=q.device)
    outs = [Q] + (list(kv_out) if kv_out is not None else [torch.empty(n, C, dtype=BF16, device=q.device) for _ in range(4)])
    _lib.call("lb_attn_prep_fwd", _p(q), _p(k), _p(kc), _p(v), _p(vc), _p(flag_sorted), _p(sorted_of), _p(pos), _p(cos_t),
              _p(sin_t), *[_p(o) for o in outs], n, heads, head_dim, _p(kv_row), _st())
    return outs   # Q, Kfv, Kfl, Vfv, Vfl


def attn_prep_fwd_bridge(q, k, v, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v, flag_sorted, sorted_of, pos, cos_t, sin_t, heads, head_dim,
                         kv_out=None, kv_row=None):
    """attn_prep_fwd with kc = k + tk.B_k^T, vc = v + tv.B_v^T computed inside the kernel (decode step: no rank-8 GEMM launch)."""
    n = q.shape[0]
    C = heads * head_dim
    rank = tk.shape[1]
    for t in (Bk_l, Bk_v, Bv_l, Bv_v):
        if tuple(t.shape) != (C, rank) or not t.is_contiguous():
            raise ValueError(f"bridge factor must be a contiguous [{C}, {rank}] tensor, got {tuple(t.shape)}")
    Q = torch.empty(n, C, dtype=BF16, device=q.device)
    outs = [Q] + (list(kv_out) if kv_out is not None else [torch.empty(n, C, dtype=BF16, device=q.device) for _ in range(4)])
    _lib.call("lb_attn_prep_fwd_bridge", _p(q), _p(k), _p(v), _p(tk), _p(tv), _p(Bk_l), _p(Bk_v), _p(Bv_l), _p(Bv_v), rank,
              _p(flag_sorted), _p(sorted_of), _p(pos), _p(cos_t), _p(sin_t), *[_p(o) for o in outs], n, heads, head_dim,
              _p(kv_row), _st())
    return outs   # Q, Kfv, Kfl, Vfv, Vfl


def attn_prep_bwd(dQ, dKfv, dKfl, dVfv, dVfl, flag_sorted, sorted_of, pos, cos_t, sin_t, heads, head_dim, bridge=True):
    n = dQ.shape[0]
    C = heads * head_dim
    dq, dk, dv = (torch.empty(n, C, dtype=BF16, device=dQ.device) for _ in range(3))
    dkb = torch.empty(n, C, dtype=BF16, device=dQ.device) if bridge else None
    dvb = torch.empty(n, C, dtype=BF16, device=dQ.device) if bridge else None
    _lib.call("lb_attn_prep_bwd", _p(dQ), _p(dKfv), _p(dKfl), _p(dVfv), _p(dVfl), _p(flag_sorted), _p(sorted_of), _p(pos),
              _p(cos_t), _p(sin_t), _p(dq), _p(dk), _p(dv), _p(dkb), _p(dvb), n, heads, head_dim, _st())
    return dq, dk, dv, dkb, dvb


_STREAM_MAX_ITEMS = None


def stream_max_cta_items() -> int:
    """Items one CTA of the persistent forward can hold (lb_attn_fwd_stream_max_cta_items)."""
    global _STREAM_MAX_ITEMS
    if _STREAM_MAX_ITEMS is None:
        _STREAM_MAX_ITEMS = int(_lib.load().lb_attn_fwd_stream_max_cta_items())
    return _STREAM_MAX_ITEMS


def attn_fwd(Q, K0, V0, K1, V1, qflag, work, kv_start, kv_end, out_row, batch, seqlen, heads, head_dim, causal, scale,
             out=None, kernel=None, plan=None):
    """kernel: "single" (two CTAs per SM, one q tile each; the default of this low-level call) or "stream" (persistent;
    `plan` = AttnWork.stream_plan(...) or None for the built-in snake split)."""
    kernel = kernel or "single"
    C = heads * head_dim
    if out is None:
        out = torch.zeros(batch * seqlen, C, dtype=BF16, device=Q.device)
    lse = torch.full((batch, heads, seqlen), float("inf"), dtype=torch.float32, device=Q.device)
    tail = (_p(kv_start), _p(kv_end), _p(out_row), _p(out), _p(lse), batch, seqlen, heads, head_dim, int(causal), float(scale), _st())
    head = (_p(Q), _p(K0), _p(V0), _p(K1), _p(V1), _p(qflag), _p(work), work.shape[0])
    if kernel == "stream":
        items, off, n_cta, max_items = plan if plan is not None else (None, None, 0, 0)
        _timed_call("lb_attn_fwd_stream", *head, _p(items), _p(off), n_cta, max_items, STREAM_HEAD_GROUP, *tail)
    else:
        _timed_call({"single": "lb_attn_fwd"}[kernel], *head, *tail)
    return out, lse


def attn_decode(Q, K_fl, V_fl, K_fv, V_fv, qflag, kv_start, kv_end, out_row, batch, heads, head_dim, kv_len, scale, out=None):
    """One new query per sample against the cached operands (lb_attn_decode).  K_*/V_*: [B, capacity, H*D]."""
    C = heads * head_dim
    capacity = K_fl.shape[1]
    if out is None:
        out = torch.empty(batch, C, dtype=BF16, device=Q.device)
    # chunks of ~256 keys, but not more CTAs than ~16 per SM (4 fit at a time: 256 threads x 64 registers)
    n_split = max(1, min((kv_len + 255) // 256, (16 * sm_count() + batch * heads - 1) // (batch * heads)))
    if os.environ.get("LB_DECODE_SPLIT"):
        n_split = max(1, int(os.environ["LB_DECODE_SPLIT"]))
    ws = torch.empty(batch * heads * n_split * (head_dim + 2), dtype=torch.float32, device=Q.device)
    _timed_call("lb_attn_decode", _p(Q), _p(K_fl), _p(V_fl), _p(K_fv), _p(V_fv), _p(qflag), _p(kv_start), _p(kv_end), _p(out_row),
                _p(ws), _p(out), batch, heads, head_dim, capacity, kv_len, n_split, float(scale), _st())
    return out


def attn_bwd_prepare(O, dO, row_of, batch, seqlen, heads, head_dim, want_dO_orig=True):
    C = heads * head_dim
    dO_orig = torch.empty(batch * seqlen, C, dtype=BF16, device=O.device) if want_dO_orig else None
    delta = torch.empty(batch, heads, seqlen, dtype=torch.float32, device=O.device)
    _lib.call("lb_attn_bwd_prepare", _p(O), _p(dO), _p(row_of), _p(dO_orig), _p(delta), batch, seqlen, heads, head_dim, _st())
    return dO_orig, delta


def attn_bwd_dq(Q, K0, V0, K1, V1, dO, lse, delta, qflag, work, kv_start, kv_end, batch, seqlen, heads, head_dim, causal,
                scale, kernel=None, plan=None):
    """kernel: "single" (one CTA per (item, head); the default of this low-level call) or "stream" (persistent, csrc/
    attn_bwd_dq_stream.cu; `plan` = AttnWork.stream_plan(...) or None for the built-in snake split)."""
    dQ = torch.empty_like(Q)          # every token row belongs to exactly one (tile, variant) work item
    head = (_p(Q), _p(K0), _p(V0), _p(K1), _p(V1), _p(dO), _p(lse), _p(delta), _p(qflag), _p(work), work.shape[0])
    tail = (_p(kv_start), _p(kv_end), _p(dQ), batch, seqlen, heads, head_dim, int(causal), float(scale), _st())
    if (kernel or "single") == "stream":
        items, off, n_cta, max_items = plan if plan is not None else (None, None, 0, 0)
        _timed_call("lb_attn_bwd_dq_stream", *head, _p(items), _p(off), n_cta, max_items, STREAM_HEAD_GROUP, *tail)
    else:
        _timed_call("lb_attn_bwd_dq", *head, *tail)
    return dQ


def attn_bwd_dkv(Q, K0, V0, K1, V1, dO, lse, delta, qflag, qtile_has, work_kv, kv_start, kv_end, batch, seqlen, heads,
                 head_dim, causal, scale, two_variants=True, kv_cover=(False, False), kernel=None, plan=None):
    """kv_cover[v]: every kv tile has a work item for variant v (host knowledge) => no zero fill needed for dK_v/dV_v.
    kernel: "single" (one CTA per (item, head); the default of this low-level call) or "stream" (persistent, csrc/
    attn_bwd_dkv_stream.cu; `plan` = AttnWork.stream_plan(..., which="kv") or None for the built-in split)."""
    mk = lambda full: torch.empty_like(K0) if full else torch.zeros_like(K0)
    dK0, dV0 = mk(kv_cover[0]), mk(kv_cover[0])
    dK1 = mk(kv_cover[1]) if two_variants else None
    dV1 = mk(kv_cover[1]) if two_variants else None
    head = (_p(Q), _p(K0), _p(V0), _p(K1), _p(V1), _p(dO), _p(lse), _p(delta), _p(qflag), _p(qtile_has), _p(work_kv), work_kv.shape[0])
    tail = (_p(kv_start), _p(kv_end), _p(dK0), _p(dV0), _p(dK1), _p(dV1), batch, seqlen, heads, head_dim, int(causal), float(scale), _st())
    if (kernel or "single") == "stream":
        items, off, n_cta, max_items = plan if plan is not None else (None, None, 0, 0)
        _timed_call("lb_attn_bwd_dkv_stream", *head, _p(items), _p(off), n_cta, max_items, STREAM_HEAD_GROUP, *tail)
    else:
        _timed_call("lb_attn_bwd_dkv", *head, *tail)
    return dK0, dV0, dK1, dV1


_DKV_STREAM = None


def dkv_stream_limits():
    """(supported, max items per CTA) of the persistent dK/dV kernel on this device."""
    global _DKV_STREAM
    if _DKV_STREAM is None:
        lib = _lib.load()
        _DKV_STREAM = (bool(lib.lb_attn_bwd_dkv_stream_supported()), int(lib.lb_attn_bwd_dkv_stream_max_cta_items()))
    return _DKV_STREAM
