"""Differentiable ops (torch.autograd.Function) over the CUDA kernels, in the decoder's *sorted-row* layout:
rows [0, n_lang) are language tokens, rows [n_lang, N) vision tokens (libra_b200.schedule.build_routing).

Plain dense GEMMs (nn.Linear-shaped products) go through cuBLAS via torch.matmul; everything else on the
hot path -- norms, SwiGLU, bridge/RoPE prologue, attention forward/backward, cross-entropy, embeddings -- is
this library's own sm_100a code.  There is no non-CUDA fallback anywhere in this module.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .schedule import AttnWork, Routing

# attention forward kernel: "stream" (csrc/attn_fwd_stream.cu: persistent, the default) or "single" (attn_fwd.cu).  A work
# list whose per-CTA share exceeds the persistent kernel's in-kernel item table runs on "single" (same work list, same
# results).
FWD_KERNEL = os.environ.get("LB_ATTN_FWD_KERNEL", "stream")


def _fwd_choice(work: AttnWork, heads: int):
    """(kernel, work list, plan) for this work list."""
    if FWD_KERNEL == "stream":
        plan = work.stream_plan(heads, ops.sm_count(), ops.STREAM_HEAD_GROUP)
        if plan[3] <= ops.stream_max_cta_items():
            return "stream", work.work_q, plan
        return "single", work.work_q, None
    return FWD_KERNEL, work.work_q, None

BF16 = torch.bfloat16


# ----------------------------------------------------------------------------- norms
class RoutedRMSNorm(torch.autograd.Function):
    """LlamaRMSNorm with the weight picked per row by modality (modeling_libra.py:463,479,817)."""

    @staticmethod
    def forward(ctx, x, w_lang, w_vis, flag, eps):
        y, rstd = ops.rmsnorm_fwd(x, w_lang, w_vis, flag, eps)
        ctx.save_for_backward(x, w_lang, w_vis, flag, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_lang, w_vis, flag, rstd = ctx.saved_tensors
        need_dw = ctx.needs_input_grad[1] or (w_vis is not None and ctx.needs_input_grad[2])
        dx, dwl, dwv = ops.rmsnorm_bwd(dy.contiguous(), x, w_lang, w_vis, flag, rstd, need_dw=need_dw)
        gl = dwl.to(w_lang.dtype) if (need_dw and ctx.needs_input_grad[1]) else None
        gv = dwv.to(w_vis.dtype) if (need_dw and w_vis is not None and ctx.needs_input_grad[2]) else None
        return dx, gl, gv, None, None


def rmsnorm(x, w_lang, w_vis=None, flag=None, eps=1e-6):
    return RoutedRMSNorm.apply(x, w_lang, w_vis, flag, eps)


class ResidualRMSNorm(torch.autograd.Function):
    """(x, norm(x)): the first output is x itself (the residual branch).  Having both consumers of x behind one node lets
    backward fold the residual gradient into the norm's backward kernel (lb_rmsnorm_bwd `residual_grad`) instead of
    autograd materialising dx_norm and adding the two gradients with a separate pass."""

    @staticmethod
    def forward(ctx, x, w_lang, w_vis, flag, eps):
        y, rstd = ops.rmsnorm_fwd(x, w_lang, w_vis, flag, eps)
        ctx.save_for_backward(x, w_lang, w_vis, flag, rstd)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, dres, dy):
        x, w_lang, w_vis, flag, rstd = ctx.saved_tensors
        need_dw = ctx.needs_input_grad[1] or (w_vis is not None and ctx.needs_input_grad[2])
        if dy is None:
            return dres, None, None, None, None
        res = None if dres is None else dres.contiguous()
        dx, dwl, dwv = ops.rmsnorm_bwd(dy.contiguous(), x, w_lang, w_vis, flag, rstd, residual_grad=res, need_dw=need_dw)
        gl = dwl.to(w_lang.dtype) if (need_dw and ctx.needs_input_grad[1]) else None
        gv = dwv.to(w_vis.dtype) if (need_dw and w_vis is not None and ctx.needs_input_grad[2]) else None
        return dx, gl, gv, None, None


def residual_rmsnorm(x, w_lang, w_vis=None, flag=None, eps=1e-6):
    """returns (x_for_residual, normed)"""
    return ResidualRMSNorm.apply(x, w_lang, w_vis, flag, eps)


class LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, mean, rstd = ops.layernorm_fwd(x, w, b, eps)
        ctx.save_for_backward(x, w, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dx, dw, db = ops.layernorm_bwd(dy.contiguous(), x, w, mean, rstd)
        return dx, dw.to(w.dtype), db.to(w.dtype), None


def layernorm(x, w, b, eps=1e-5):
    return LayerNorm.apply(x, w, b, eps)


# ----------------------------------------------------------------------------- activations
class SwiGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, up):
        ctx.save_for_backward(gate, up)
        return ops.swiglu_fwd(gate, up)

    @staticmethod
    def backward(ctx, dout):
        gate, up = ctx.saved_tensors
        dg, du = ops.swiglu_bwd(dout.contiguous(), gate, up)
        return dg, du


def swiglu(gate, up):
    return SwiGLU.apply(gate, up)


class BiasQuickGelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias):
        ctx.save_for_backward(x, bias)
        return ops.bias_quick_gelu_fwd(x, bias)

    @staticmethod
    def backward(ctx, dy):
        x, bias = ctx.saved_tensors
        dx = ops.bias_quick_gelu_bwd(dy.contiguous(), x, bias)
        db = dx.reshape(-1, dx.shape[-1]).sum(0).to(bias.dtype) if (bias is not None and ctx.needs_input_grad[1]) else None
        return dx, db


def bias_quick_gelu(x, bias):
    return BiasQuickGelu.apply(x, bias)


# ----------------------------------------------------------------------------- two-stream routing
# The language GEMMs (M = n_lang rows, N = 4096..11008) and the vision low-rank GEMMs (M = n_vis rows, N = 1024/2752 then
# 4096/11008) of one routed projection are independent.  The vision ones are small: x_v A^T has only ~40 output tiles of
# 256x256 for 148 SMs, so on one stream they leave most of the machine idle (and the language GEMM's last wave is partial
# too).  Issuing the vision path on a side stream lets the hardware co-schedule the two tile sets; fork/join are events.
USE_SIDE_STREAM = True
_side = {}


def _side_stream():
    dev = torch.cuda.current_device()
    st = _side.get(dev)
    if st is None:
        st = _side[dev] = torch.cuda.Stream(device=dev)
    return st


class _Fork:
    """with _Fork() as f:  ...main-stream work...;  with f.side(): ...side-stream work...   (join on exit)"""

    def __init__(self, enabled=True):
        self.on = bool(enabled and USE_SIDE_STREAM and torch.cuda.is_available())

    def __enter__(self):
        if self.on:
            self.main = torch.cuda.current_stream()
            self.st = _side_stream()
            self.st.wait_event(self.main.record_event())
        return self

    def side(self):
        return torch.cuda.stream(self.st) if self.on else _Null()

    def __exit__(self, *exc):
        if self.on:
            self.main.wait_event(self.st.record_event())
        return False


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


# ----------------------------------------------------------------------------- routed linear
# Weight-gradient accumulation fusion: when a parameter already owns a .grad buffer (e.g. a view into the flat
# data-parallel gradient buffer, libra_b200.dist.FlatGradBuffer), dW is accumulated straight into it by the GEMM
# (beta = 1) instead of being materialised and added by autograd afterwards (saves ~3 passes over the weight size).
FUSE_WGRAD_ACCUMULATE = True


def _wgrad(W: torch.Tensor, a_t: torch.Tensor, b: torch.Tensor):
    """dW = a_t @ b, either returned (autograd accumulates) or accumulated in place into W.grad (returns None)."""
    g = W.grad
    if FUSE_WGRAD_ACCUMULATE and g is not None and g.dtype == a_t.dtype and g.is_contiguous():
        g.addmm_(a_t, b)
        return None
    return torch.matmul(a_t, b)


class RoutedLinear(torch.autograd.Function):
    """y[:n_lang] = x[:n_lang] W^T ;  y[n_lang:] = (x[n_lang:] A^T) B^T   (+ residual, fused into the GEMM as beta = 1)
    (language nn.Linear | vision LibraLinear, modeling_libra.py:192-199, routed by :129-147).
    The two row ranges are contiguous, so there is no gather/scatter and no boolean indexing."""

    @staticmethod
    def forward(ctx, x, n_lang, W, A, B, residual):
        N = x.shape[0]
        y = torch.empty(N, W.shape[0], dtype=x.dtype, device=x.device)
        mid = torch.empty(N - n_lang, A.shape[0], dtype=x.dtype, device=x.device) if N - n_lang > 0 else None
        with _Fork(n_lang > 0 and N - n_lang > 0) as f:
            if N - n_lang > 0:
                with f.side():
                    torch.matmul(x[n_lang:], A.t(), out=mid)
                    if residual is None:
                        torch.matmul(mid, B.t(), out=y[n_lang:])
                    else:
                        torch.addmm(residual[n_lang:], mid, B.t(), out=y[n_lang:])
            if n_lang > 0:
                if residual is None:
                    torch.matmul(x[:n_lang], W.t(), out=y[:n_lang])
                else:
                    torch.addmm(residual[:n_lang], x[:n_lang], W.t(), out=y[:n_lang])
        ctx.n_lang = n_lang
        ctx.save_for_backward(x, W, A, B, mid)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, A, B, mid = ctx.saved_tensors
        n_lang = ctx.n_lang
        N = x.shape[0]
        dy = dy.contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dW = dA = dB = None
        dmid = torch.empty_like(mid) if mid is not None else None
        fusable = lambda P_: FUSE_WGRAD_ACCUMULATE and P_.grad is not None and P_.grad.dtype == dy.dtype
        # non-fused weight gradients are allocated here, on the main stream, before any side-stream work
        if N - n_lang > 0:
            if ctx.needs_input_grad[4] and not fusable(B):
                dB = torch.empty_like(B)
            if ctx.needs_input_grad[3] and not fusable(A):
                dA = torch.empty_like(A)
        with _Fork(n_lang > 0 and N - n_lang > 0) as f:
            if N - n_lang > 0:
                with f.side():
                    dyv = dy[n_lang:]
                    torch.matmul(dyv, B, out=dmid)
                    if dx is not None:
                        torch.matmul(dmid, A, out=dx[n_lang:])
                    if ctx.needs_input_grad[4]:
                        if dB is None:
                            B.grad.addmm_(dyv.t(), mid)
                        else:
                            torch.matmul(dyv.t(), mid, out=dB)
                    if ctx.needs_input_grad[3]:
                        if dA is None:
                            A.grad.addmm_(dmid.t(), x[n_lang:])
                        else:
                            torch.matmul(dmid.t(), x[n_lang:], out=dA)
            if n_lang > 0:
                if dx is not None:
                    torch.matmul(dy[:n_lang], W, out=dx[:n_lang])
                if ctx.needs_input_grad[2]:
                    dW = _wgrad(W, dy[:n_lang].t(), x[:n_lang])
        if n_lang == 0 and ctx.needs_input_grad[2]:
            dW = torch.zeros_like(W)
        if N - n_lang == 0:
            if ctx.needs_input_grad[3]:
                dA = torch.zeros_like(A)
            if ctx.needs_input_grad[4]:
                dB = torch.zeros_like(B)
        return dx, None, dW, dA, dB, (dy if ctx.needs_input_grad[5] else None)


def routed_linear(x, n_lang, W, A, B, residual=None):
    return RoutedLinear.apply(x, n_lang, W, A, B, residual)


class RoutedFanout(torch.autograd.Function):
    """Several routed projections of the SAME input in one autograd node (q/k/v + the two bridge down-projections, or
    gate + up).  Forward is the same GEMMs as RoutedLinear / RoutedDown; backward accumulates every branch's input
    gradient into one buffer through the GEMM (beta = 1), replacing autograd's N-1 full-size gradient additions.
    Language GEMMs run on the current stream, vision GEMMs on the side stream (see _Fork).
    kinds[i] == "lin":  weights (W, A, B) -> y = [x_l W^T ; (x_v A^T) B^T];   "down": weights (A_lang, A_vis)."""

    @staticmethod
    def forward(ctx, x, n_lang, kinds, *weights):
        N = x.shape[0]
        nv = N - n_lang
        xl, xv = x[:n_lang], x[n_lang:]
        outs, mids, specs, wi = [], [], [], 0
        for kind in kinds:                      # allocate everything on the main stream first
            if kind == "lin":
                W, A, B = weights[wi:wi + 3]
                wi += 3
                y = torch.empty(N, W.shape[0], dtype=x.dtype, device=x.device)
                mid = torch.empty(nv, A.shape[0], dtype=x.dtype, device=x.device) if nv > 0 else None
                specs.append((kind, y, mid, W, A, B))
            else:
                Al, Av = weights[wi:wi + 2]
                wi += 2
                y = torch.empty(N, Al.shape[0], dtype=x.dtype, device=x.device)
                mid = None
                specs.append((kind, y, None, Al, Av, None))
            outs.append(y)
            mids.append(mid)
        with _Fork(n_lang > 0 and nv > 0) as f:
            if nv > 0:
                with f.side():
                    for kind, y, mid, W0, W1, W2 in specs:
                        if kind == "lin":
                            torch.matmul(xv, W1.t(), out=mid)
                            torch.matmul(mid, W2.t(), out=y[n_lang:])
                        else:
                            torch.matmul(xv, W1.t(), out=y[n_lang:])
            if n_lang > 0:
                for kind, y, mid, W0, W1, W2 in specs:
                    torch.matmul(xl, W0.t(), out=y[:n_lang])
        ctx.n_lang, ctx.kinds = n_lang, kinds
        ctx.mid_present = [m is not None for m in mids]
        ctx.save_for_backward(x, *weights, *[m for m in mids if m is not None])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        saved = ctx.saved_tensors
        x = saved[0]
        nw = sum(3 if k == "lin" else 2 for k in ctx.kinds)
        weights = saved[1:1 + nw]
        mids_saved = list(saved[1 + nw:])
        n, N = ctx.n_lang, x.shape[0]
        nv = N - n
        ng = ctx.needs_input_grad
        need_dx = ng[0]
        dx = torch.empty_like(x) if need_dx else None
        fusable = lambda P_: FUSE_WGRAD_ACCUMULATE and P_.grad is not None and P_.grad.dtype == x.dtype
        # plan (and allocate on the main stream) per branch
        plan, grads, wi = [], [], 0
        for oi, kind in enumerate(ctx.kinds):
            dy = douts[oi]
            dy = dy.contiguous() if dy is not None else None
            if kind == "lin":
                W, A, B = weights[wi:wi + 3]
                mid = mids_saved.pop(0) if ctx.mid_present[oi] else None
                e = dict(kind=kind, dy=dy, W=W, A=A, B=B, mid=mid, gi=len(grads),
                         needW=ng[3 + wi], needA=ng[3 + wi + 1], needB=ng[3 + wi + 2])
                if dy is not None and nv > 0:
                    e["dmid"] = torch.empty_like(mid)
                    e["gA"] = torch.empty_like(A) if (e["needA"] and not fusable(A)) else None
                    e["gB"] = torch.empty_like(B) if (e["needB"] and not fusable(B)) else None
                grads += [None, None, None]
                wi += 3
            else:
                Al, Av = weights[wi:wi + 2]
                e = dict(kind=kind, dy=dy, Al=Al, Av=Av, gi=len(grads), needL=ng[3 + wi], needV=ng[3 + wi + 1])
                if dy is not None and nv > 0:
                    e["gV"] = torch.empty_like(Av) if (e["needV"] and not fusable(Av)) else None
                grads += [None, None]
                wi += 2
            plan.append(e)

        def acc(dst, a, b, first):
            if first:
                torch.matmul(a, b, out=dst)
            else:
                dst.addmm_(a, b)

        first_l = first_v = True
        with _Fork(n > 0 and nv > 0) as f:
            if nv > 0:
                with f.side():
                    for e in plan:
                        dy = e["dy"]
                        if dy is None:
                            continue
                        if e["kind"] == "lin":
                            torch.matmul(dy[n:], e["B"], out=e["dmid"])
                            if need_dx:
                                acc(dx[n:], e["dmid"], e["A"], first_v)
                                first_v = False
                            if e["needB"]:
                                if e["gB"] is None:
                                    e["B"].grad.addmm_(dy[n:].t(), e["mid"])
                                else:
                                    torch.matmul(dy[n:].t(), e["mid"], out=e["gB"])
                                    grads[e["gi"] + 2] = e["gB"]
                            if e["needA"]:
                                if e["gA"] is None:
                                    e["A"].grad.addmm_(e["dmid"].t(), x[n:])
                                else:
                                    torch.matmul(e["dmid"].t(), x[n:], out=e["gA"])
                                    grads[e["gi"] + 1] = e["gA"]
                        else:
                            if need_dx:
                                acc(dx[n:], dy[n:], e["Av"], first_v)
                                first_v = False
                            if e["needV"]:
                                if e["gV"] is None:
                                    e["Av"].grad.addmm_(dy[n:].t(), x[n:])
                                else:
                                    torch.matmul(dy[n:].t(), x[n:], out=e["gV"])
                                    grads[e["gi"] + 1] = e["gV"]
            if n > 0:
                for e in plan:
                    dy = e["dy"]
                    if dy is None:
                        continue
                    if e["kind"] == "lin":
                        if need_dx:
                            acc(dx[:n], dy[:n], e["W"], first_l)
                            first_l = False
                        if e["needW"]:
                            grads[e["gi"]] = _wgrad(e["W"], dy[:n].t(), x[:n])
                    else:
                        if need_dx:
                            acc(dx[:n], dy[:n], e["Al"], first_l)
                            first_l = False
                        if e["needL"]:
                            grads[e["gi"]] = _wgrad(e["Al"], dy[:n].t(), x[:n])
        if need_dx:
            if first_l and n > 0:
                dx[:n].zero_()
            if first_v and nv > 0:
                dx[n:].zero_()
        return (dx, None, None, *grads)


def routed_fanout(x, n_lang, kinds, *weights):
    return RoutedFanout.apply(x, n_lang, tuple(kinds), *weights)


class RoutedDown(torch.autograd.Function):
    """t[:n_lang] = x[:n_lang] A_lang^T ; t[n_lang:] = x[n_lang:] A_vis^T  -- the rank-r first halves of the
    bridge LibraLinears (vision_{k,v}_bridge_on_{language,vision}.weight_A, modeling_libra.py:259-263,318-319)."""

    @staticmethod
    def forward(ctx, x, n_lang, A_lang, A_vis):
        N = x.shape[0]
        t = torch.empty(N, A_lang.shape[0], dtype=x.dtype, device=x.device)
        if n_lang > 0:
            torch.matmul(x[:n_lang], A_lang.t(), out=t[:n_lang])
        if N - n_lang > 0:
            torch.matmul(x[n_lang:], A_vis.t(), out=t[n_lang:])
        ctx.n_lang = n_lang
        ctx.save_for_backward(x, A_lang, A_vis)
        return t

    @staticmethod
    def backward(ctx, dt):
        x, A_lang, A_vis = ctx.saved_tensors
        n = ctx.n_lang
        dt = dt.contiguous()
        dx = torch.empty_like(x)
        torch.matmul(dt[:n], A_lang, out=dx[:n])
        torch.matmul(dt[n:], A_vis, out=dx[n:])
        dAl = torch.matmul(dt[:n].t(), x[:n]) if ctx.needs_input_grad[2] else None
        dAv = torch.matmul(dt[n:].t(), x[n:]) if ctx.needs_input_grad[3] else None
        return dx, None, dAl, dAv


def routed_down(x, n_lang, A_lang, A_vis):
    return RoutedDown.apply(x, n_lang, A_lang, A_vis)


# ----------------------------------------------------------------------------- attention
@dataclass
class AttnMeta:
    routing: Routing
    work: AttnWork
    pos: torch.Tensor            # [B*T] int32 rotary position per original token
    cos: torch.Tensor            # [n_pos, D/2] fp32
    sin: torch.Tensor
    batch: int
    seqlen: int
    heads: int
    head_dim: int
    # decode path (N1): the cache the layers append their key/value operands to, the layer in flight, and -- for a one-token
    # step -- the visible key range per sample (int32 [B] or None)
    kv_cache: Optional[object] = None
    layer_idx: int = 0
    decode: bool = False
    dec_kv_start: Optional[torch.Tensor] = None
    dec_kv_end: Optional[torch.Tensor] = None
    dec_kv_row: Optional[torch.Tensor] = None          # int32 [B]: b * capacity + length, the new token's row of the cache


class BridgeAttention(torch.autograd.Function):
    """LibraAttention core with use_bridge=True (modeling_libra.py:318-397, 267-296): bridge add + RoPE prologue,
    tcgen05 flash attention over the two key/value variants, output scattered back to sorted rows."""

    @staticmethod
    def forward(ctx, q, k, v, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v, meta: AttnMeta):
        rt, w = meta.routing, meta.work
        n = rt.n_lang
        # bridged variants k + tk.Bk^T, v + tv.Bv^T per modality segment: rank-r GEMMs with beta = 1 (cuBLAS)
        kc, vc = torch.empty_like(k), torch.empty_like(v)
        if n > 0:
            torch.addmm(k[:n], tk[:n], Bk_l.t(), out=kc[:n])
            torch.addmm(v[:n], tv[:n], Bv_l.t(), out=vc[:n])
        if rt.n_vis > 0:
            torch.addmm(k[n:], tk[n:], Bk_v.t(), out=kc[n:])
            torch.addmm(v[n:], tv[n:], Bv_v.t(), out=vc[n:])
        scale = 1.0 / math.sqrt(meta.head_dim)
        o = torch.empty_like(q)
        if meta.decode:
            # one-token step (N1): the prologue writes the new key/value operands straight into the token's cache slot
            cache, i = meta.kv_cache, meta.layer_idx
            Q = ops.attn_prep_fwd(q, k, kc, v, vc, rt.flag_sorted, rt.inv, meta.pos, meta.cos, meta.sin, meta.heads, meta.head_dim,
                                  kv_out=(cache.k_fv[i], cache.k_fl[i], cache.v_fv[i], cache.v_fl[i]), kv_row=meta.dec_kv_row)[0]
        else:
            Q, Kfv, Kfl, Vfv, Vfl = ops.attn_prep_fwd(q, k, kc, v, vc, rt.flag_sorted, rt.inv, meta.pos, meta.cos, meta.sin,
                                                     meta.heads, meta.head_dim)
        del kc, vc
        if meta.kv_cache is not None:                      # use_cache=True (modeling_libra.py:343-361): inference only
            cache = meta.kv_cache
            if not meta.decode:
                cache.append(meta.layer_idx, Kfv, Kfl, Vfv, Vfl, meta.seqlen)
            if meta.decode:
                i = meta.layer_idx
                # with a device-side key range the host-side length only sizes the split: keep it static (graph replay)
                kv_len = cache.capacity if meta.dec_kv_end is not None else cache.length + 1
                ops.attn_decode(Q, cache.k_fl[i], cache.v_fl[i], cache.k_fv[i], cache.v_fv[i], rt.flag_orig, meta.dec_kv_start,
                                meta.dec_kv_end, rt.inv, meta.batch, meta.heads, meta.head_dim, kv_len, scale, out=o)
                return o
        # variant 0 = language queries (see Kfl/Vfl), variant 1 = vision queries (see Kfv/Vfv); rows land in sorted order
        kern, wlist, plan = _fwd_choice(w, meta.heads)
        o, lse = ops.attn_fwd(Q, Kfl, Vfl, Kfv, Vfv, rt.flag_orig, wlist, w.kv_start,
                              w.kv_end, rt.inv, meta.batch, meta.seqlen, meta.heads, meta.head_dim, True, scale, out=o,
                              kernel=kern, plan=plan)
        ctx.meta = meta
        ctx.scale = scale
        ctx.save_for_backward(Q, Kfv, Kfl, Vfv, Vfl, o, lse, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v)
        return o

    @staticmethod
    def backward(ctx, do):
        meta: AttnMeta = ctx.meta
        rt, w = meta.routing, meta.work
        Q, Kfv, Kfl, Vfv, Vfl, o, lse, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v = ctx.saved_tensors
        B, T, H, D = meta.batch, meta.seqlen, meta.heads, meta.head_dim
        dO, delta = ops.attn_bwd_prepare(o, do.contiguous(), rt.inv, B, T, H, D)
        dQ = ops.attn_bwd_dq(Q, Kfl, Vfl, Kfv, Vfv, dO, lse, delta, rt.flag_orig, w.work_q, w.kv_start, w.kv_end, B, T, H, D,
                             True, ctx.scale)
        dKfl, dVfl, dKfv, dVfv = ops.attn_bwd_dkv(Q, Kfl, Vfl, Kfv, Vfv, dO, lse, delta, rt.flag_orig, w.qtile_has, w.work_kv,
                                                 w.kv_start, w.kv_end, B, T, H, D, True, ctx.scale, kv_cover=w.kv_cover)
        dq, dk, dv, dkb, dvb = ops.attn_prep_bwd(dQ, dKfv, dKfl, dVfv, dVfl, rt.flag_sorted, rt.inv, meta.pos, meta.cos,
                                                 meta.sin, H, D)
        n = rt.n_lang
        # kb = tk . B^T  =>  d_tk = dkb . B ; dB = dkb^T . tk   (per modality segment)
        d_tk = torch.empty_like(tk)
        d_tv = torch.empty_like(tv)
        torch.matmul(dkb[:n], Bk_l, out=d_tk[:n])
        torch.matmul(dkb[n:], Bk_v, out=d_tk[n:])
        torch.matmul(dvb[:n], Bv_l, out=d_tv[:n])
        torch.matmul(dvb[n:], Bv_v, out=d_tv[n:])
        g = ctx.needs_input_grad
        dBk_l = torch.matmul(dkb[:n].t(), tk[:n]) if g[5] else None
        dBk_v = torch.matmul(dkb[n:].t(), tk[n:]) if g[6] else None
        dBv_l = torch.matmul(dvb[:n].t(), tv[:n]) if g[7] else None
        dBv_v = torch.matmul(dvb[n:].t(), tv[n:]) if g[8] else None
        return dq, dk, dv, d_tk, d_tv, dBk_l, dBk_v, dBv_l, dBv_v, None


def bridge_attention(q, k, v, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v, meta: AttnMeta):
    return BridgeAttention.apply(q, k, v, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v, meta)


class PlainAttention(torch.autograd.Function):
    """Non-causal self-attention over [B*T, H*D] (CLIPAttention core, modeling_clip.py:309-349).
    `scale` multiplies q.k^T inside the kernel (the reference pre-multiplies q by head_dim**-0.5, :299)."""

    @staticmethod
    def forward(ctx, q, k, v, work: AttnWork, batch, seqlen, heads, head_dim, scale):
        kern, wlist, plan = _fwd_choice(work, heads)
        o, lse = ops.attn_fwd(q, k, v, None, None, None, wlist, None, None,
                              None, batch, seqlen, heads, head_dim, False, scale, kernel=kern, plan=plan)
        ctx.args = (work, batch, seqlen, heads, head_dim, scale)
        ctx.save_for_backward(q, k, v, o, lse)
        return o

    @staticmethod
    def backward(ctx, do):
        work, B, T, H, D, scale = ctx.args
        q, k, v, o, lse = ctx.saved_tensors
        dO, delta = ops.attn_bwd_prepare(o, do.contiguous(), None, B, T, H, D, want_dO_orig=False)
        dO = do.contiguous()
        dq = ops.attn_bwd_dq(q, k, v, None, None, dO, lse, delta, None, work.work_q, None, None, B, T, H, D, False, scale)
        dk, dv, _, _ = ops.attn_bwd_dkv(q, k, v, None, None, dO, lse, delta, None, work.qtile_has, work.work_kv, None, None, B,
                                        T, H, D, False, scale, two_variants=False)
        return dq, dk, dv, None, None, None, None, None, None


def plain_attention(q, k, v, work, batch, seqlen, heads, head_dim, scale=1.0):
    return PlainAttention.apply(q.contiguous(), k.contiguous(), v.contiguous(), work, batch, seqlen, heads, head_dim, scale)


# ----------------------------------------------------------------------------- embeddings
class EmbedLang(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table):
        ctx.save_for_backward(ids)
        ctx.shape = table.shape
        return ops.embed_lang(ids, table)

    @staticmethod
    def backward(ctx, dy):
        (ids,) = ctx.saved_tensors
        dt = torch.zeros(ctx.shape, dtype=torch.float32, device=dy.device)
        dy = dy.contiguous()
        ops.embed_bwd(ids, dy, 0, ctx.shape[1], dt)
        return None, dt.to(dy.dtype)


class EmbedVisionCat(torch.autograd.Function):
    """[vemb0[id0] | vemb1[id1] | signal] per vision row (modeling_libra.py:629-644)."""

    @staticmethod
    def forward(ctx, ids0, ids1, table0, table1, signal, signal_row, signal_cols):
        ctx.save_for_backward(ids0, ids1)
        ctx.shapes = (table0.shape, table1.shape)
        return ops.embed_vision_cat(ids0, ids1, table0, table1, signal, signal_row, signal_cols)

    @staticmethod
    def backward(ctx, dy):
        ids0, ids1 = ctx.saved_tensors
        s0, s1 = ctx.shapes
        dy = dy.contiguous()
        d0 = torch.zeros(s0, dtype=torch.float32, device=dy.device)
        d1 = torch.zeros(s1, dtype=torch.float32, device=dy.device)
        ops.embed_bwd(ids0, dy, 0, s0[1], d0)
        ops.embed_bwd(ids1, dy, s0[1], s1[1], d1)
        return None, None, d0.to(dy.dtype), d1.to(dy.dtype), None, None, None


# ----------------------------------------------------------------------------- heads + loss
class HeadCrossEntropy(torch.autograd.Function):
    """sum over rows of CE(x W^T, labels) * row_scale, fused: logits are produced by one GEMM, reduced and turned
    into their own gradient in place by lb_cross_entropy_fwd_bwd, and never leave bf16 / HBM more than 3 times
    (modeling_libra.py:1018-1052 restricted to the finite vocabulary block of the row's modality, :1159-1174).
    Returns (loss_sum fp32 scalar tensor, n_valid fp32 scalar tensor)."""

    @staticmethod
    def forward(ctx, x, W, labels, grad_scale):
        logits = torch.matmul(x, W.t())
        row_loss = ops.cross_entropy_fwd_bwd(logits, labels, W.shape[0], grad_scale)
        ctx.save_for_backward(x, W, logits)
        return row_loss.sum()

    @staticmethod
    def backward(ctx, g):
        x, W, dlogits = ctx.saved_tensors
        dx = torch.matmul(dlogits, W) if ctx.needs_input_grad[0] else None
        dW = torch.matmul(dlogits.t(), x) if ctx.needs_input_grad[1] else None      # scaled by g below: not fusable
        # upstream gradient of the (already pre-scaled) partial loss is a scalar
        if dx is not None:
            dx = dx * g.to(dx.dtype)
        if dW is not None:
            dW = dW * g.to(dW.dtype)
        return dx, dW, None, None


def head_cross_entropy(x, W, labels, grad_scale=1.0):
    return HeadCrossEntropy.apply(x, W, labels, grad_scale)
