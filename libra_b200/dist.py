"""Data-parallel plumbing: every rank holds the full bf16 weights and runs the hot path on its
own samples; the only exchange is ONE all-reduce over a flat gradient buffer per optimizer step (NCCL over
NVLink 5 / NVSwitch on the GPU box; gloo in the CPU tests).  Replaces the reference's DeepSpeed ZeRO-2 reduce-scatter +
all-gather (libra/configs/deepspeed_configs/ZeRO-2.json:16-19) at the level of averaged gradients."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradBuffer:
    """Owns one contiguous gradient buffer; every trainable parameter's .grad is a view into it, so backward
    accumulates in place and the reduction is a single collective on a single message."""

    def __init__(self, params: Iterable[torch.nn.Parameter], dtype: Optional[torch.dtype] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        dtype = dtype or self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dtype, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None, chunks: int = 1):
        """sum over ranks / world size.  chunks > 1 issues the same buffer as a few large contiguous pieces (in
        reverse-layer order of creation) so a caller can overlap them with the tail of backward."""
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        world = dist.get_world_size(group)
        if world == 1:
            return self.flat
        if chunks <= 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        else:
            step = (self.numel + chunks - 1) // chunks
            for i in reversed(range(chunks)):
                dist.all_reduce(self.flat[i * step:(i + 1) * step], op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(world)
        return self.flat


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
