"""Differentiable ops (torch.autograd.Function) over the CUDA kernels, in the decoder's *sorted-row* layout:
rows [0, n_lang) are language tokens, rows [n_lang, N) vision tokens (libra_b200.schedule.build_routing).

Everything on the hot path -- every dense product (grouped persistent tcgen05 GEMM, csrc/gemm_grouped.cu), norms,
SwiGLU, bridge/RoPE prologue, attention forward/backward, cross-entropy, embeddings -- is this library's own sm_100a
code; nothing here calls cuBLAS.  There is no non-CUDA fallback anywhere in this module.
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

# dQ kernel of the attention backward: "stream" (csrc/attn_bwd_dq_stream.cu, persistent; shares the forward's plan) or "single"
DQ_KERNEL = os.environ.get("LB_ATTN_DQ_KERNEL", "stream")


def _dq_choice(work: AttnWork, heads: int):
    if DQ_KERNEL == "stream":
        plan = work.stream_plan(heads, ops.sm_count(), ops.STREAM_HEAD_GROUP)
        if plan[3] <= ops.stream_max_cta_items():
            return "stream", plan
    return "single", None


# dK/dV kernel of the attention backward: "stream" (csrc/attn_bwd_dkv_stream.cu, persistent) or "single"
DKV_KERNEL = os.environ.get("LB_ATTN_DKV_KERNEL", "stream")


def _dkv_choice(work: AttnWork, heads: int):
    if DKV_KERNEL == "stream" and work.kv_tiles is not None:
        ok, max_items = ops.dkv_stream_limits()
        if ok:
            sms = ops.sm_count()
            n_items = len(work.kv_tiles) * heads
            waves = max(1, -(-n_items // (sms * max(1, max_items - 8))))          # CTAs beyond one per SM queue up behind the first wave
            plan = work.stream_plan(heads, sms * waves, ops.STREAM_HEAD_GROUP, which="kv")
            if plan[3] <= max_items:
                return "stream", plan
    return "single", None


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
        gl = _param_grad(w_lang, dwl.to(w_lang.dtype)) if (need_dw and ctx.needs_input_grad[1]) else None
        gv = _param_grad(w_vis, dwv.to(w_vis.dtype)) if (need_dw and w_vis is not None and ctx.needs_input_grad[2]) else None
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
        gl = _param_grad(w_lang, dwl.to(w_lang.dtype)) if (need_dw and ctx.needs_input_grad[1]) else None
        gv = _param_grad(w_vis, dwv.to(w_vis.dtype)) if (need_dw and w_vis is not None and ctx.needs_input_grad[2]) else None
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


# ----------------------------------------------------------------------------- dense products
# Every nn.Linear-shaped product of the decoder runs on this library's grouped persistent tcgen05 kernel
# (csrc/gemm_grouped.cu, ops.gemm_grouped): one launch per routed projection holds the dense language problem, the
# chained low-rank vision problems (mid = x_v A^T, y_v = mid B^T, the second waiting on the first per row block) and, in
# backward, every dgrad / wgrad product of the node.  Input gradients of a fan-out are ONE problem whose K loop runs
# over the branches (no beta = 1 passes over dx); weight gradients accumulate straight into a parameter's gradient
# buffer when the owner of that buffer asked for it (see mark_fused_grad).
G = ops.gp


def mark_fused_grad(params, fresh: bool = True):
    """Opt parameters in to fused weight-gradient accumulation: backward then adds dW into the existing `.grad` buffer
    inside the GEMM epilogue and returns None to autograd for that parameter.  Only the owner of the gradient buffers
    may do this (libra_b200.dist.FlatGradBuffer, bench.py): no AccumulateGrad hook fires for such a parameter, so it is
    incompatible with hook-based gradient synchronisation (DDP, FSDP) and with torch.autograd.grad -- which is why it is
    off unless asked for.  fresh=True: the buffers hold no gradient yet (start of an optimizer step); the first
    backward OVERWRITES them (beta = 0), which replaces zeroing the buffer."""
    for p_ in params:
        p_._lb_fused_grad = True
        p_._lb_grad_fresh = bool(fresh)


def begin_grad_step(params):
    """Start of an optimizer step for fused-grad parameters: the next weight gradient overwrites instead of adding."""
    for p_ in params:
        if getattr(p_, "_lb_fused_grad", False):
            p_._lb_grad_fresh = True


def _fused(P_: torch.Tensor, dtype) -> bool:
    g = P_.grad
    return bool(getattr(P_, "_lb_fused_grad", False)) and g is not None and g.dtype == dtype and g.is_contiguous()


def _param_grad(P_, g):
    """Hand a materialised parameter gradient to autograd, or (fused-grad parameters) put it into the buffer here."""
    if g is None or not _fused(P_, P_.dtype):
        return g
    if getattr(P_, "_lb_grad_fresh", False):
        P_._lb_grad_fresh = False
        P_.grad.copy_(g)
    else:
        P_.grad.add_(g.to(P_.grad.dtype))
    return None


def _no_grad_contribution(P_):
    """A needed parameter gradient that is identically zero (empty modality segment, unused branch)."""
    if _fused(P_, P_.dtype):
        if getattr(P_, "_lb_grad_fresh", False):
            P_._lb_grad_fresh = False
            P_.grad.zero_()
        return None
    return torch.zeros_like(P_)


def _wgrad_entry(P_, a, b, need, **kw):
    """(entry, grad to hand to autograd): dP = a^T . b  with a: [rows, out], b: [rows, in].  None, None if not needed."""
    if not need:
        return None, None
    if _fused(P_, a.dtype):
        g = P_.grad
        if getattr(P_, "_lb_grad_fresh", False):
            P_._lb_grad_fresh = False
            return G(a, b, g, ta=True, tb=True, **kw), None
        return G(a, b, g, ta=True, tb=True, d=g, **kw), None
    g = torch.empty_like(P_)
    return G(a, b, g, ta=True, tb=True, **kw), g


class _Launch:
    """Collects the entries of one grouped launch; entry indices are what `wait_on` refers to."""

    def __init__(self):
        self.entries = []

    def add(self, e):
        if e is None:
            return -1
        self.entries.append(e)
        return len(self.entries) - 1

    def run(self):
        es = self.entries
        # the kernel takes 32 entries per launch; more (never on the decoder's shapes) would need a split that keeps
        # wait_on / acc_prev groups together
        if len(es) > 32:
            raise RuntimeError(f"grouped GEMM launch with {len(es)} entries")
        ops.gemm_grouped(es)


class RoutedLinear(torch.autograd.Function):
    """y[:n_lang] = x[:n_lang] W^T ;  y[n_lang:] = (x[n_lang:] A^T) B^T   (+ residual, added in the GEMM epilogue)
    (language nn.Linear | vision LibraLinear, modeling_libra.py:192-199, routed by :129-147).
    The two row ranges are contiguous, so there is no gather/scatter and no boolean indexing; one launch."""

    @staticmethod
    def forward(ctx, x, n_lang, W, A, B, residual):
        N = x.shape[0]
        nv = N - n_lang
        y = torch.empty(N, W.shape[0], dtype=x.dtype, device=x.device)
        mid = torch.empty(nv, A.shape[0], dtype=x.dtype, device=x.device) if nv > 0 else None
        L = _Launch()
        if nv > 0:
            i1 = L.add(G(x[n_lang:], A, mid))
        if n_lang > 0:
            L.add(G(x[:n_lang], W, y[:n_lang], d=None if residual is None else residual[:n_lang]))
        if nv > 0:
            L.add(G(mid, B, y[n_lang:], d=None if residual is None else residual[n_lang:], wait_on=i1))
        L.run()
        ctx.n_lang = n_lang
        ctx.save_for_backward(x, W, A, B, mid)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, A, B, mid = ctx.saved_tensors
        n = ctx.n_lang
        N = x.shape[0]
        nv = N - n
        dy = dy.contiguous()
        ng = ctx.needs_input_grad
        dx = torch.empty_like(x) if ng[0] else None
        dW = dA = dB = None
        L = _Launch()
        if nv > 0:
            dmid = torch.empty_like(mid)
            i1 = L.add(G(dy[n:], B, dmid, tb=True))
        if n > 0:
            if dx is not None:
                L.add(G(dy[:n], W, dx[:n], tb=True))
            e, dW = _wgrad_entry(W, dy[:n], x[:n], ng[2])
            L.add(e)
        elif ng[2]:
            dW = _no_grad_contribution(W)
        if nv > 0:
            if dx is not None:
                L.add(G(dmid, A, dx[n:], tb=True, wait_on=i1))
            e, dB = _wgrad_entry(B, dy[n:], mid, ng[4])
            L.add(e)
            e, dA = _wgrad_entry(A, dmid, x[n:], ng[3], wait_on=i1)
            L.add(e)
        else:
            dA = _no_grad_contribution(A) if ng[3] else None
            dB = _no_grad_contribution(B) if ng[4] else None
        L.run()
        return dx, None, dW, dA, dB, (dy if ng[5] else None)


def routed_linear(x, n_lang, W, A, B, residual=None):
    return RoutedLinear.apply(x, n_lang, W, A, B, residual)


def _fanout_forward(x, n_lang, kinds, weights, swiglu_pair=None):
    """The forward launch of a fan-out.  Returns (outs, mids).  swiglu_pair = (i, j, act, keep): branches i (gate) and j (up)
    additionally produce act = silu(gate) * up for the language rows in the same tile pass (SwiGLU epilogue); keep = the
    pre-activations of the language rows are stored too (backward needs them; inference does not -- and without them a
    one-token decode step's gate|up product is a plain problem the weight-streaming kernel takes, ops.gemm_skinny)."""
    N = x.shape[0]
    nv = N - n_lang
    xl, xv = x[:n_lang], x[n_lang:]
    outs, mids, specs, wi = [], [], [], 0
    for kind in kinds:
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
    L = _Launch()
    stage1 = {}
    if nv > 0:                                   # producers of the chains first: their dependents come last
        for bi, (kind, y, mid, W0, W1, W2) in enumerate(specs):
            if kind == "lin":
                stage1[bi] = L.add(G(xv, W1, mid))
        for kind, y, mid, W0, W1, W2 in specs:
            if kind == "down":
                L.add(G(xv, W1, y[n_lang:]))
    if n_lang > 0:
        skip = set()
        if swiglu_pair is not None:
            gi, ui, act, keep = swiglu_pair
            L.add(G(xl, specs[gi][3], act[:n_lang], b2=specs[ui][3], epi=ops.EPI_SWIGLU, g=specs[gi][1][:n_lang] if keep else None,
                    u=specs[ui][1][:n_lang] if keep else None))
            skip = {gi, ui}
        for bi, (kind, y, mid, W0, W1, W2) in enumerate(specs):
            if bi not in skip:
                L.add(G(xl, W0, y[:n_lang]))
    if nv > 0:
        for bi, (kind, y, mid, W0, W1, W2) in enumerate(specs):
            if kind == "lin":
                L.add(G(mid, W2, y[n_lang:], wait_on=stage1[bi]))
    L.run()
    return outs, mids


def _fanout_backward(x, n, kinds, weights, mids, douts, ng_x, ng_w):
    """One launch: dx (one K-segmented problem per modality), every weight gradient, the chains' dmid."""
    N = x.shape[0]
    nv = N - n
    dx = torch.empty_like(x) if ng_x else None
    grads = []
    L = _Launch()
    plan, wi = [], 0
    for oi, kind in enumerate(kinds):
        dy = douts[oi]
        dy = dy.contiguous() if dy is not None else None
        if kind == "lin":
            W, A, B = weights[wi:wi + 3]
            plan.append(dict(kind=kind, dy=dy, W=W, A=A, B=B, mid=mids[oi], need=ng_w[wi:wi + 3], gi=len(grads)))
            grads += [None, None, None]
            wi += 3
        else:
            Al, Av = weights[wi:wi + 2]
            plan.append(dict(kind=kind, dy=dy, Al=Al, Av=Av, need=ng_w[wi:wi + 2], gi=len(grads)))
            grads += [None, None]
            wi += 2
    live = [e for e in plan if e["dy"] is not None]
    # ---- vision: dmid producers first
    if nv > 0:
        for e in live:
            if e["kind"] == "lin":
                e["dmid"] = torch.empty_like(e["mid"])
                e["i_dmid"] = L.add(G(e["dy"][n:], e["B"], e["dmid"], tb=True))
    # ---- language
    if n > 0:
        if dx is not None:
            first = True
            for e in live:
                Wl = e["W"] if e["kind"] == "lin" else e["Al"]
                L.add(G(e["dy"][:n], Wl, dx[:n], tb=True, acc_prev=not first))
                first = False
        for e in live:
            Wl = e["W"] if e["kind"] == "lin" else e["Al"]
            ent, g = _wgrad_entry(Wl, e["dy"][:n], x[:n], e["need"][0])
            L.add(ent)
            grads[e["gi"]] = g
    # ---- vision dependents
    if nv > 0:
        if dx is not None:
            first = True
            for e in live:
                if e["kind"] == "lin":
                    L.add(G(e["dmid"], e["A"], dx[n:], tb=True, wait_on=e["i_dmid"], acc_prev=not first))
                else:
                    L.add(G(e["dy"][n:], e["Av"], dx[n:], tb=True, acc_prev=not first))
                first = False
        for e in live:
            if e["kind"] == "lin":
                ent, g = _wgrad_entry(e["B"], e["dy"][n:], e["mid"], e["need"][2])
                L.add(ent)
                grads[e["gi"] + 2] = g
                ent, g = _wgrad_entry(e["A"], e["dmid"], x[n:], e["need"][1], wait_on=e["i_dmid"])
                L.add(ent)
                grads[e["gi"] + 1] = g
            else:
                ent, g = _wgrad_entry(e["Av"], e["dy"][n:], x[n:], e["need"][1])
                L.add(ent)
                grads[e["gi"] + 1] = g
    L.run()
    if dx is not None and not live:
        dx.zero_()
    # parameters of an empty modality segment / of branches without an incoming gradient
    wi = 0
    for e in plan:
        k = 3 if e["kind"] == "lin" else 2
        for j in range(k):
            if ng_w[wi + j] and grads[e["gi"] + j] is None:
                seg_empty = (n == 0) if j == 0 else (nv == 0)
                if e["dy"] is None or seg_empty:
                    grads[e["gi"] + j] = _no_grad_contribution(weights[wi + j])
        wi += k
    return dx, grads


class RoutedFanout(torch.autograd.Function):
    """Several routed projections of the SAME input in one autograd node and one launch per direction (q/k/v + the two
    bridge down-projections).  kinds[i] == "lin": weights (W, A, B) -> y = [x_l W^T ; (x_v A^T) B^T];
    "down": weights (A_lang, A_vis) -> [x_l A_lang^T ; x_v A_vis^T]."""

    @staticmethod
    def forward(ctx, x, n_lang, kinds, *weights):
        outs, mids = _fanout_forward(x, n_lang, kinds, weights)
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
        ms = list(saved[1 + nw:])
        mids = [ms.pop(0) if pres else None for pres in ctx.mid_present]
        ng = ctx.needs_input_grad
        dx, grads = _fanout_backward(x, ctx.n_lang, ctx.kinds, weights, mids, douts, ng[0], ng[3:])
        return (dx, None, None, *grads)


def routed_fanout(x, n_lang, kinds, *weights):
    return RoutedFanout.apply(x, n_lang, tuple(kinds), *weights)


class RoutedGateUp(torch.autograd.Function):
    """act = silu(gate(x)) * up(x) with routed gate / up projections (LibraMLP.forward, modeling_libra.py:227-238, the
    product at :232-233).  Language rows: ONE problem whose B tile interleaves the gate and up weights, the SwiGLU product
    taken in the epilogue (gate / up pre-activations are stored for backward by the same tile pass); vision rows: the two
    low-rank chains in the same launch, then lb_swiglu_fwd on those rows."""

    @staticmethod
    def forward(ctx, x, n_lang, keep, Wg, Ag, Bg, Wu, Au, Bu):
        N = x.shape[0]
        act = torch.empty(N, Wg.shape[0], dtype=x.dtype, device=x.device)
        weights = (Wg, Ag, Bg, Wu, Au, Bu)
        (g, u), mids = _fanout_forward(x, n_lang, ("lin", "lin"), weights, swiglu_pair=(0, 1, act, keep))
        if N - n_lang > 0:
            ops.swiglu_fwd(g[n_lang:], u[n_lang:], out=act[n_lang:])
        ctx.n_lang = n_lang
        ctx.mid_present = [m is not None for m in mids]
        ctx.save_for_backward(x, g, u, *weights, *[m for m in mids if m is not None])
        return act

    @staticmethod
    def backward(ctx, dact):
        saved = ctx.saved_tensors
        x, g, u = saved[:3]
        weights = saved[3:9]
        ms = list(saved[9:])
        mids = [ms.pop(0) if pres else None for pres in ctx.mid_present]
        dg, du = ops.swiglu_bwd(dact.contiguous(), g, u)
        ng = ctx.needs_input_grad
        dx, grads = _fanout_backward(x, ctx.n_lang, ("lin", "lin"), weights, mids, (dg, du), ng[0], ng[3:])
        return (dx, None, None, *grads)


def routed_gate_up(x, n_lang, Wg, Ag, Bg, Wu, Au, Bu):
    # the language rows' gate / up pre-activations are written only when a backward pass can follow (inside forward() the grad
    # mode is always off, so the decision is taken here)
    keep = torch.is_grad_enabled() and any(t.requires_grad for t in (x, Wg, Ag, Bg, Wu, Au, Bu))
    return RoutedGateUp.apply(x, n_lang, keep, Wg, Ag, Bg, Wu, Au, Bu)


class Linear(torch.autograd.Function):
    """y = x W^T (+ bias) (+ quick_gelu) (+ residual): nn.Linear on the grouped kernel (CLIP projections, modeling_clip.py:
    279-282, 371-378; the vision signal projection, modeling_libra.py:640-645).  act: 0 none, 1 quick_gelu."""

    @staticmethod
    def forward(ctx, x, W, bias, residual, act):
        x2 = x.reshape(-1, x.shape[-1])
        Nout = W.shape[0]
        ld = (Nout + 7) // 8 * 8                   # TMA needs a 16-byte row pitch; odd widths (18, 514) get a padded buffer
        y = torch.empty(x2.shape[0], ld, dtype=x.dtype, device=x.device)[:, :Nout]
        if bias is not None and Nout % 8:          # the epilogue reads the bias in 8-element vectors
            bias = torch.cat([bias, bias.new_zeros(ld - Nout)])
        pre = None
        need_pre = act == 1 and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        if need_pre:
            pre = torch.empty_like(y)
        res2 = None if residual is None else residual.reshape(-1, Nout)
        if act == 1:
            ops.gemm_grouped([G(x2, W, y, bias=bias, epi=ops.EPI_QGELU, g=pre)])
            if res2 is not None:
                y = y + res2
        else:
            ops.gemm_grouped([G(x2, W, y, bias=bias, d=res2)])
        ctx.act = act
        ctx.xshape = x.shape
        ctx.save_for_backward(x2, W, pre)
        return y.view(*x.shape[:-1], Nout) if ld == Nout else y.reshape(*x.shape[:-1], Nout)

    @staticmethod
    def backward(ctx, dy):
        x2, W, pre = ctx.saved_tensors
        ng = ctx.needs_input_grad
        dy2 = dy.reshape(-1, W.shape[0]).contiguous()
        dres = dy if ng[3] else None
        if ctx.act == 1:
            dy2 = ops.bias_quick_gelu_bwd(dy2, pre, None)
        dx = dW = db = None
        es = []
        if ng[0]:
            dx = torch.empty_like(x2)
            es.append(G(dy2, W, dx, tb=True))
        if ng[1]:
            e, dW = _wgrad_entry(W, dy2, x2, True)
            es.append(e)
        ops.gemm_grouped(es)
        if ng[2]:
            db = dy2.sum(0).to(W.dtype)
        return (None if dx is None else dx.view(ctx.xshape)), dW, db, dres, None


def linear(x, W, bias=None, residual=None, act=0):
    return Linear.apply(x, W, bias, residual, act)


class LinearFanout(torch.autograd.Function):
    """Several biased nn.Linear projections of the same input in one launch (CLIP q/k/v, modeling_clip.py:295-301);
    backward sums the input gradients in one K-segmented problem."""

    @staticmethod
    def forward(ctx, x, *wb):
        x2 = x.reshape(-1, x.shape[-1])
        Ws, bs = wb[0::2], wb[1::2]
        ys = [torch.empty(x2.shape[0], W.shape[0], dtype=x.dtype, device=x.device) for W in Ws]
        ops.gemm_grouped([G(x2, W, y, bias=b) for W, b, y in zip(Ws, bs, ys)])
        ctx.xshape = x.shape
        ctx.nb = len(Ws)
        ctx.save_for_backward(x2, *Ws)
        return tuple(y.view(*x.shape[:-1], y.shape[1]) for y in ys)

    @staticmethod
    def backward(ctx, *dys):
        x2 = ctx.saved_tensors[0]
        Ws = ctx.saved_tensors[1:]
        ng = ctx.needs_input_grad
        dys = [dy.reshape(-1, W.shape[0]).contiguous() for dy, W in zip(dys, Ws)]
        es, out = [], [None] * (2 * ctx.nb)
        dx = None
        if ng[0]:
            dx = torch.empty_like(x2)
            for i, (dy, W) in enumerate(zip(dys, Ws)):
                es.append(G(dy, W, dx, tb=True, acc_prev=i > 0))
        for i, (dy, W) in enumerate(zip(dys, Ws)):
            if ng[1 + 2 * i]:
                e, out[2 * i] = _wgrad_entry(W, dy, x2, True)
                es.append(e)
            if ng[2 + 2 * i]:
                out[2 * i + 1] = dy.sum(0).to(W.dtype)
        ops.gemm_grouped(es)
        return ((None if dx is None else dx.view(ctx.xshape)), *out)


def linear_fanout(x, *weights_and_biases):
    return LinearFanout.apply(x, *weights_and_biases)


class RoutedDown(torch.autograd.Function):
    """t[:n_lang] = x[:n_lang] A_lang^T ; t[n_lang:] = x[n_lang:] A_vis^T  -- the rank-r first halves of the
    bridge LibraLinears (vision_{k,v}_bridge_on_{language,vision}.weight_A, modeling_libra.py:259-263,318-319)."""

    @staticmethod
    def forward(ctx, x, n_lang, A_lang, A_vis):
        (t,), _ = _fanout_forward(x, n_lang, ("down",), (A_lang, A_vis))
        ctx.n_lang = n_lang
        ctx.save_for_backward(x, A_lang, A_vis)
        return t

    @staticmethod
    def backward(ctx, dt):
        x, A_lang, A_vis = ctx.saved_tensors
        ng = ctx.needs_input_grad
        dx, grads = _fanout_backward(x, ctx.n_lang, ("down",), (A_lang, A_vis), [None], (dt,), ng[0], ng[2:])
        return (dx, None, *grads)


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


# Decode step: compute the rank-8 bridge products inside the attention prologue instead of a separate skinny-GEMM launch.
# Measured on Libra-11B (B=8, graph replay, dependent launches on): 5.51-5.55 ms per step without, 5.60-5.62 ms with -- the
# prologue grows from 4.4 to 9.3 us on its 8 CTAs while the 8.5 us GEMM launch it replaces already overlapped its neighbours.
# Off by default; LB_FOLD_DECODE_BRIDGE=1 selects it (tests cover both).
FOLD_DECODE_BRIDGE = os.environ.get("LB_FOLD_DECODE_BRIDGE", "0") == "1"


class BridgeAttention(torch.autograd.Function):
    """LibraAttention core with use_bridge=True (modeling_libra.py:318-397, 267-296): bridge add + RoPE prologue,
    tcgen05 flash attention over the two key/value variants, output scattered back to sorted rows."""

    @staticmethod
    def forward(ctx, q, k, v, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v, meta: AttnMeta):
        rt, w = meta.routing, meta.work
        n = rt.n_lang
        scale = 1.0 / math.sqrt(meta.head_dim)
        o = torch.empty_like(q)
        kc = vc = None
        fold = meta.decode and FOLD_DECODE_BRIDGE and tk.shape[1] % 8 == 0
        if not fold:
            # bridged variants k + tk.Bk^T, v + tv.Bv^T per modality segment: four rank-r problems with the addend in the
            # epilogue, one launch
            kc, vc = torch.empty_like(k), torch.empty_like(v)
            es = []
            if n > 0:
                es += [G(tk[:n], Bk_l, kc[:n], d=k[:n]), G(tv[:n], Bv_l, vc[:n], d=v[:n])]
            if rt.n_vis > 0:
                es += [G(tk[n:], Bk_v, kc[n:], d=k[n:]), G(tv[n:], Bv_v, vc[n:], d=v[n:])]
            ops.gemm_grouped(es)
        if meta.decode:
            # one-token step (N1): the prologue writes the new key/value operands straight into the token's cache slot
            # (optionally computing the rank-r bridge products itself: FOLD_DECODE_BRIDGE)
            cache, i = meta.kv_cache, meta.layer_idx
            kv_out = (cache.k_fv[i], cache.k_fl[i], cache.v_fv[i], cache.v_fl[i])
            if fold:
                Q = ops.attn_prep_fwd_bridge(q, k, v, tk, tv, Bk_l, Bk_v, Bv_l, Bv_v, rt.flag_sorted, rt.inv, meta.pos, meta.cos,
                                             meta.sin, meta.heads, meta.head_dim, kv_out=kv_out, kv_row=meta.dec_kv_row)[0]
            else:
                Q = ops.attn_prep_fwd(q, k, kc, v, vc, rt.flag_sorted, rt.inv, meta.pos, meta.cos, meta.sin, meta.heads,
                                      meta.head_dim, kv_out=kv_out, kv_row=meta.dec_kv_row)[0]
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
        dq_kern, dq_plan = _dq_choice(w, H)
        dQ = ops.attn_bwd_dq(Q, Kfl, Vfl, Kfv, Vfv, dO, lse, delta, rt.flag_orig, w.work_q, w.kv_start, w.kv_end, B, T, H, D,
                             True, ctx.scale, kernel=dq_kern, plan=dq_plan)
        dkv_kern, dkv_plan = _dkv_choice(w, H)
        dKfl, dVfl, dKfv, dVfv = ops.attn_bwd_dkv(Q, Kfl, Vfl, Kfv, Vfv, dO, lse, delta, rt.flag_orig, w.qtile_has, w.work_kv,
                                                 w.kv_start, w.kv_end, B, T, H, D, True, ctx.scale, kv_cover=w.kv_cover,
                                                 kernel=dkv_kern, plan=dkv_plan)
        dq, dk, dv, dkb, dvb = ops.attn_prep_bwd(dQ, dKfv, dKfl, dVfv, dVfl, rt.flag_sorted, rt.inv, meta.pos, meta.cos,
                                                 meta.sin, H, D)
        n = rt.n_lang
        # kb = tk . B^T  =>  d_tk = dkb . B ; dB = dkb^T . tk   (per modality segment), one launch
        d_tk = torch.empty_like(tk)
        d_tv = torch.empty_like(tv)
        g = ctx.needs_input_grad
        es, dB = [], [None] * 4
        segs = ((slice(0, n), Bk_l, Bv_l, 0), (slice(n, None), Bk_v, Bv_v, 1))
        for sl, Bk, Bv, vi in segs:
            if (n if vi == 0 else rt.n_vis) == 0:
                dB[vi] = _no_grad_contribution(Bk) if g[5 + vi] else None
                dB[2 + vi] = _no_grad_contribution(Bv) if g[7 + vi] else None
                continue
            es += [G(dkb[sl], Bk, d_tk[sl], tb=True), G(dvb[sl], Bv, d_tv[sl], tb=True)]
            e, dB[vi] = _wgrad_entry(Bk, dkb[sl], tk[sl], g[5 + vi])
            es.append(e)
            e, dB[2 + vi] = _wgrad_entry(Bv, dvb[sl], tv[sl], g[7 + vi])
            es.append(e)
        ops.gemm_grouped([e for e in es if e is not None])
        dBk_l, dBk_v, dBv_l, dBv_v = dB
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
        dq_kern, dq_plan = _dq_choice(work, H)
        dq = ops.attn_bwd_dq(q, k, v, None, None, dO, lse, delta, None, work.work_q, None, None, B, T, H, D, False, scale,
                             kernel=dq_kern, plan=dq_plan)
        dkv_kern, dkv_plan = _dkv_choice(work, H)
        dk, dv, _, _ = ops.attn_bwd_dkv(q, k, v, None, None, dO, lse, delta, None, work.qtile_has, work.work_kv, None, None, B,
                                        T, H, D, False, scale, two_variants=False, kernel=dkv_kern, plan=dkv_plan)
        return dq, dk, dv, None, None, None, None, None, None


def plain_attention(q, k, v, work, batch, seqlen, heads, head_dim, scale=1.0):
    return PlainAttention.apply(q.contiguous(), k.contiguous(), v.contiguous(), work, batch, seqlen, heads, head_dim, scale)


# ----------------------------------------------------------------------------- embeddings
class EmbedLang(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table, padding_idx=None):
        ctx.save_for_backward(ids)
        ctx.shape = table.shape
        ctx.table = table                       # the parameter object (its .grad buffer may take the gradient directly)
        ctx.padding_idx = padding_idx
        return ops.embed_lang(ids, table)

    @staticmethod
    def backward(ctx, dy):
        (ids,) = ctx.saved_tensors
        dt = torch.zeros(ctx.shape, dtype=torch.float32, device=dy.device)
        dy = dy.contiguous()
        ops.embed_bwd(ids, dy, 0, ctx.shape[1], dt)
        if ctx.padding_idx is not None:         # nn.Embedding(padding_idx=...) never updates that row (modeling_libra.py:539)
            dt[ctx.padding_idx].zero_()
        return None, _param_grad(ctx.table, dt.to(dy.dtype)), None


class EmbedVisionCat(torch.autograd.Function):
    """[vemb0[id0] | vemb1[id1] | signal] per vision row (modeling_libra.py:629-644)."""

    @staticmethod
    def forward(ctx, ids0, ids1, table0, table1, signal, signal_row, signal_cols):
        ctx.save_for_backward(ids0, ids1)
        ctx.shapes = (table0.shape, table1.shape)
        ctx.tables = (table0, table1)
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
        return None, None, _param_grad(ctx.tables[0], d0.to(dy.dtype)), _param_grad(ctx.tables[1], d1.to(dy.dtype)), None, None, None


# ----------------------------------------------------------------------------- heads + loss
class HeadCrossEntropy(torch.autograd.Function):
    """sum over rows of CE(x W^T, labels) * row_scale, fused: logits are produced by one GEMM, reduced and turned
    into their own gradient in place by lb_cross_entropy_fwd_bwd, and never leave bf16 / HBM more than 3 times
    (modeling_libra.py:1018-1052 restricted to the finite vocabulary block of the row's modality, :1159-1174).
    The upstream (scalar) gradient multiplies the two backward products inside their epilogues (alpha)."""

    @staticmethod
    def forward(ctx, x, W, labels, grad_scale):
        V = W.shape[0]
        ld = (V + 7) // 8 * 8                       # TMA rows: a 16-byte multiple (the 514-wide vision heads pad to 520)
        buf = torch.empty(x.shape[0], ld, dtype=x.dtype, device=x.device)
        logits = buf[:, :V]
        ops.gemm_grouped([G(x, W, logits)])
        row_loss = ops.cross_entropy_fwd_bwd(logits, labels, V, grad_scale)
        ctx.save_for_backward(x, W, buf)
        return row_loss.sum()

    @staticmethod
    def backward(ctx, g):
        x, W, buf = ctx.saved_tensors
        dlogits = buf[:, :W.shape[0]]
        alpha = g.detach().to(torch.float32).reshape(1)
        es, dx, dW = [], None, None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            es.append(G(dlogits, W, dx, tb=True, alpha=alpha))
        e, dW = _wgrad_entry(W, dlogits, x, ctx.needs_input_grad[1], alpha=alpha)
        if e is not None:
            es.append(e)
        ops.gemm_grouped(es)
        return dx, dW, None, None


def head_cross_entropy(x, W, labels, grad_scale=1.0):
    return HeadCrossEntropy.apply(x, W, labels, grad_scale)
