"""Data-parallel plumbing: every rank holds the full bf16 weights and runs the hot path on its
own samples; the only exchange is ONE all-reduce over a flat gradient buffer per optimizer step (NCCL over
NVLink 5 / NVSwitch on the GPU box; gloo in the CPU tests).  Replaces the reference's DeepSpeed ZeRO-2 reduce-scatter +
all-gather (libra/configs/deepspeed_configs/ZeRO-2.json:16-19) at the level of averaged gradients.

The buffer is laid out in the order gradients become FINAL during backward -- heads and final norms, then decoder layer
L-1 down to layer 0, then the embedding side -- so "everything ready so far" is always a prefix.  `GradSync` issues that
prefix in a few large pieces while the last micro-batch's backward is still running (the pieces are still one logical
all-reduce of one buffer); what cannot overlap is the last piece: the embedding-side parameters (< 2 % of Libra-11B).
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def readiness_order(named_params: Sequence[Tuple[str, torch.nn.Parameter]]) -> List[Tuple[str, torch.nn.Parameter, int]]:
    """(name, param, group) sorted by the moment the gradient is final in backward.
    group 0: head side (final norms, lm_head, vision_lm_head, anything that is not a layer or an embedding);
    group 1 + (L-1-i): decoder layer i;  last group: embedding side (token / vision embeddings, vision signal path)."""
    layer_re = re.compile(r"(?:^|\.)layers\.(\d+)\.")
    items = []
    n_layers = 0
    for n, p in named_params:
        m = layer_re.search(n)
        if m:
            n_layers = max(n_layers, int(m.group(1)) + 1)
    for n, p in named_params:
        m = layer_re.search(n)
        if m:
            g = 1 + (n_layers - 1 - int(m.group(1)))
        elif "embed_tokens" in n or "vision_contiguous_signal_processor" in n or "vision_signal_norm" in n:
            g = n_layers + 1
        else:
            g = 0
        items.append((n, p, g))
    items.sort(key=lambda t: t[2])               # stable: registration order inside a group
    return items


class FlatGradBuffer:
    """Owns one contiguous gradient buffer (and optionally one contiguous weight buffer); every trainable parameter's
    .grad (and .data) is a view into it, so backward accumulates in place, the reduction is a single message and the
    optimizer a handful of launches.  The parameters are opted in to fused weight-gradient accumulation
    (functional.mark_fused_grad): the GEMM epilogues write dW straight into the views, the first micro-batch of a step
    overwriting instead of adding, so the buffer is never zeroed."""

    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], dtype: Optional[torch.dtype] = None,
                 flatten_weights: bool = False, fused: bool = True):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        if not named:
            raise ValueError("no trainable parameters")
        order = readiness_order(named)
        self.names = [n for n, _, _ in order]
        self.params: List[torch.nn.Parameter] = [p for _, p, _ in order]
        self.groups = [g for _, _, g in order]
        dev = self.params[0].device
        dtype = dtype or self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dtype, device=dev)
        self.flat_w = torch.empty(self.numel, dtype=self.params[0].dtype, device=dev) if flatten_weights else None
        self.offsets: Dict[str, Tuple[int, int]] = {}
        self.group_end: Dict[int, int] = {}
        off = 0
        with torch.no_grad():
            for n, p, g in order:
                k = p.numel()
                if self.flat_w is not None:
                    self.flat_w[off:off + k].copy_(p.reshape(-1))
                    p.data = self.flat_w[off:off + k].view_as(p)
                p.grad = self.flat[off:off + k].view_as(p)
                self.offsets[n] = (off, off + k)
                off += k
                self.group_end[g] = off
        self.fused = bool(fused) and self.flat.is_cuda
        if self.fused:
            from . import functional as LF
            LF.mark_fused_grad(self.params, fresh=True)

    def begin_step(self):
        """Start of an optimizer step: the first backward overwrites the (stale) gradients.  Without fused accumulation
        (CPU tests, foreign autograd functions) the buffer is zeroed instead."""
        if self.fused:
            from . import functional as LF
            LF.begin_grad_step(self.params)
        else:
            self.flat.zero_()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        """Unoverlapped fallback: one all-reduce of the whole buffer, then / world."""
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        world = dist.get_world_size(group)
        if world == 1:
            return self.flat
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(world)
        return self.flat


class GradSync:
    """Overlapped gradient reduction for one model replica per rank.

        sync = GradSync(buf, model.model)                 # LibraModel exposes `layer_grad_ready_hook`
        for step:
            buf.begin_step()
            for i, micro in enumerate(micros):
                sync.arm(last=(i == len(micros) - 1))
                loss(micro).backward()
            sync.finish()                                  # issues the rest, waits; buffer holds the SUM over ranks
    The loss of every rank is pre-divided by the world size by the caller (mean gradient), as bench.py does.
    `min_bytes`: a piece is issued once at least that much of the ready prefix is pending (launch latency vs overlap)."""

    def __init__(self, buf: FlatGradBuffer, hook_owner=None, group=None, min_bytes: int = 1 << 30, n_layers: Optional[int] = None):
        self.buf, self.group = buf, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.min_elems = max(1, min_bytes // buf.flat.element_size())
        self.n_layers = n_layers if n_layers is not None else (max(buf.groups) - 1)
        self.issued = 0
        self.pending = []
        self.armed = False
        self.pieces: List[Tuple[int, int]] = []          # (lo, hi) of every piece of the last step: tests / traces
        if hook_owner is not None:
            hook_owner.layer_grad_ready_hook = self.on_layer_grad_ready

    def arm(self, last: bool):
        self.armed = bool(last) and self.world > 1
        if last:
            self.issued = 0
            self.pieces = []

    def _issue(self, hi: int):
        if hi <= self.issued:
            return
        lo = self.issued
        self.pending.append(dist.all_reduce(self.buf.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.pieces.append((lo, hi))
        self.issued = hi

    def on_layer_grad_ready(self, li: int):
        """d(loss)/d(input of layer li) exists => the gradients of the head side and of layers >= li are final."""
        if not self.armed:
            return
        g = 1 + (self.n_layers - 1 - li)
        hi = self.buf.group_end.get(g)
        if hi is None:
            return
        if hi - self.issued >= self.min_elems or li == 0:
            self._issue(hi)

    def finish(self):
        """After the last backward: reduce what is left (the embedding side), wait for every piece."""
        if self.world > 1:
            self._issue(self.buf.numel)
            for w in self.pending:
                w.wait()
        self.pending.clear()
        self.armed = False
        return self.buf.flat


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Split a global batch evenly by sample across ranks (the only partitioning the path needs)."""
    out = {}
    for k, v in batch.items():
        if v is None:
            out[k] = None
            continue
        bdim = 1 if k in ("input_ids", "labels") else 0       # [Q,B,T] planes carry the batch in dim 1
        n = v.shape[bdim]
        assert n % world == 0, f"{k}: batch {n} not divisible by world size {world}"
        out[k] = v.narrow(bdim, rank * (n // world), n // world).contiguous()
    return out
