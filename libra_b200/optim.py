"""Optimizer of the reference's training recipe on flat bf16 buffers.

Reference policy (trainer.py:27-85, libra/configs/libra_pretrain.yaml:81-91,116): AdamW(beta 0.9 / 0.99, eps 1e-8),
weight decay 0.01 on every parameter that is not inside a LayerNorm / LlamaRMSNorm module and whose name has no "bias"
(`get_decay_parameter_names`), gradient clipping at max_grad_norm 1.0, cosine schedule with warm-up ratio 0.05.
Here: ONE fused kernel launch over the flat buffers (lb_adamw_bf16_scaled; the decay exclusions are a device-side range table),
the global gradient norm and the clip factor computed on the device (lb_grad_clip_scale: one read of the gradient buffer)
and folded into that same pass -- no host synchronisation and no separate scaling pass over the 22 GB gradient buffer.
"""
from __future__ import annotations

import ctypes
import math
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .ops import _p, _st


def decay_parameter_names(model: torch.nn.Module, norm_types: Optional[Tuple[type, ...]] = None) -> List[str]:
    """LibraTrainer.get_decay_parameter_names (trainer.py:27-37): parameters outside LayerNorm / LlamaRMSNorm modules whose
    name does not contain "bias"."""
    if norm_types is None:
        from .models.modeling_libra import LlamaRMSNorm
        norm_types = (torch.nn.LayerNorm, LlamaRMSNorm)
    skip = set()
    for mn, mod in model.named_modules():
        if isinstance(mod, norm_types):
            for pn, _ in mod.named_parameters(recurse=False):
                skip.add(f"{mn}.{pn}" if mn else pn)
    return [n for n, _ in model.named_parameters() if n not in skip and "bias" not in n]


def cosine_with_warmup(step: int, total_steps: int, warmup_steps: int) -> float:
    """transformers.get_cosine_schedule_with_warmup factor (lr_scheduler_type "cosine", libra_pretrain.yaml:82)."""
    if step < warmup_steps:
        return step / max(1, warmup_steps)
    prog = (step - warmup_steps) / max(1, total_steps - warmup_steps)
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * min(1.0, prog))))


class FlatAdamW:
    """AdamW (torch.optim.AdamW semantics) over flat bf16 parameter / gradient buffers whose slices are the model's
    parameters (libra_b200.dist.FlatGradBuffer).  States are bf16 like the parameters, matching what `model.to(bf16)` +
    torch AdamW gives in the reference recipe (train.py:31-32).

    runs: [(lo, hi, weight_decay)] covering the buffer (None: one run with `weight_decay`); max_grad_norm > 0 clips the
    global gradient norm inside the update pass; schedule(step) -> lr factor."""

    def __init__(self, flat_param: torch.Tensor, flat_grad: torch.Tensor, lr=1e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0,
                 runs: Optional[Sequence[Tuple[int, int, float]]] = None, max_grad_norm: float = 0.0, schedule=None):
        assert flat_param.dtype == torch.bfloat16 and flat_grad.dtype == torch.bfloat16
        assert flat_param.numel() == flat_grad.numel()
        self.p, self.g = flat_param, flat_grad
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(flat_param)
        self.v = torch.zeros_like(flat_param)
        self.t = 0
        n = flat_param.numel()
        self.runs = self._align(list(runs) if runs is not None else [(0, n, float(weight_decay))], n)
        self.max_grad_norm = float(max_grad_norm)
        self.schedule = schedule
        self.clip_ws = torch.empty(2368, dtype=torch.float32, device=flat_param.device) if self.max_grad_norm > 0 else None
        self.clip_out = torch.ones(2, dtype=torch.float32, device=flat_param.device)      # [norm, scale]
        self.last_lr = lr

    @staticmethod
    def _align(runs, n):
        """Merge neighbours of equal decay and move run boundaries to multiples of 8 elements (the kernel's vector width; every
        parameter of the model has a multiple of 8 elements, so this only guards odd test shapes -- the remainder goes to
        `tail`)."""
        runs = sorted(runs)
        out = []
        for lo, hi, wd in runs:
            if out and out[-1][2] == wd and out[-1][1] == lo:
                out[-1] = (out[-1][0], hi, wd)
            else:
                out.append((lo, hi, wd))
        assert out[0][0] == 0 and out[-1][1] == n and all(a[1] == b[0] for a, b in zip(out, out[1:])), "runs must tile the buffer"
        return out

    @classmethod
    def for_buffer(cls, buf, model: torch.nn.Module, lr=1e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0,
                   total_steps: Optional[int] = None, warmup_ratio: float = 0.05):
        """The reference recipe on a FlatGradBuffer built with flatten_weights=True."""
        assert buf.flat_w is not None, "FlatGradBuffer(flatten_weights=True) required"
        decay = set(decay_parameter_names(model))
        runs = [(lo, hi, weight_decay if n in decay else 0.0) for n, (lo, hi) in sorted(buf.offsets.items(), key=lambda kv: kv[1][0])]
        sched = None
        if total_steps:
            warm = int(math.ceil(total_steps * warmup_ratio))
            sched = lambda s: cosine_with_warmup(s, total_steps, warm)
        return cls(buf.flat_w, buf.flat, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, runs=runs,
                   max_grad_norm=max_grad_norm, schedule=sched)

    def grad_norm(self) -> torch.Tensor:
        """device tensor [norm, clip factor] of the last step() (no sync until read)."""
        return self.clip_out

    def _plan(self):
        """One launch over the 8-aligned body of the buffer with weight decay `wd_main`, the no-decay runs as a device range
        table; what does not fit that form (other decay values, < 8 stray elements at unaligned run edges) goes to `tails`."""
        if getattr(self, "_planned", None) is not None:
            return self._planned
        n = self.p.numel()
        decays = sorted({wd for _, _, wd in self.runs if wd != 0.0})
        wd_main = decays[0] if decays else 0.0
        ranges, tails = [], []
        for lo, hi, wd in self.runs:
            lo8, hi8 = (lo + 7) // 8 * 8, hi // 8 * 8
            if wd == 0.0 and wd_main != 0.0 and hi8 > lo8:
                ranges.append((lo8 // 8, hi8 // 8))
            if wd not in (0.0, wd_main):
                raise NotImplementedError("FlatAdamW: one non-zero weight decay value per buffer")
        # elements of a run that are not 8-aligned inside their run share a vector with the neighbour run: do them on the side
        for (lo, hi, wd), nxt in zip(self.runs, self.runs[1:] + [None]):
            if hi % 8 and nxt is not None and nxt[2] != wd:
                v0 = hi // 8 * 8
                tails.append((slice(v0, min(v0 + 8, n)),))
        n8 = n - n % 8
        if n % 8:
            tails.append((slice(n8, n),))
        tbl = torch.tensor(ranges, dtype=torch.int64, device=self.p.device).reshape(-1, 2) if ranges else None
        self._planned = (wd_main, tbl, tails, n8)
        return self._planned

    def step(self):
        self.t += 1
        lr = self.lr * (self.schedule(self.t - 1) if self.schedule is not None else 1.0)      # HF steps the scheduler after the update
        self.last_lr = lr
        wd_main, tbl, tails, n8 = self._plan()
        scale = None
        if self.max_grad_norm > 0:
            _lib.call("lb_grad_clip_scale", _p(self.g), n8, float(self.max_grad_norm), _p(self.clip_ws), self.clip_ws.numel(),
                      _p(self.clip_out), _st())
            scale = ctypes.c_void_p(self.clip_out.data_ptr() + 4)
        saved = [(sl, self.p[sl].clone(), self.m[sl].clone(), self.v[sl].clone()) for (sl,) in tails]
        _lib.call("lb_adamw_bf16_scaled", _p(self.p), _p(self.g), _p(self.m), _p(self.v), n8, float(lr), float(self.betas[0]),
                  float(self.betas[1]), float(self.eps), float(wd_main), self.t, scale, _p(tbl), 0 if tbl is None else tbl.shape[0],
                  _st())
        for sl, p0, m0, v0 in saved:                      # shared / stray vectors: redo element-wise with each element's own decay
            self.p[sl], self.m[sl], self.v[sl] = p0, m0, v0
            for i in range(sl.start, sl.stop):
                self._tail(slice(i, i + 1), lr, self._wd_of(i))

    def _wd_of(self, i):
        for lo, hi, wd in self.runs:
            if lo <= i < hi:
                return wd
        return 0.0

    def _tail(self, sl, lr, wd):
        gs = self.clip_out[1] if self.max_grad_norm > 0 else 1.0
        p, g = self.p[sl].float(), self.g[sl].float() * gs
        m = self.m[sl].float().mul_(self.betas[0]).add_(g, alpha=1 - self.betas[0])
        v = self.v[sl].float().mul_(self.betas[1]).addcmul_(g, g, value=1 - self.betas[1])
        bc1, bc2 = 1 - self.betas[0] ** self.t, 1 - self.betas[1] ** self.t
        p.mul_(1 - lr * wd).addcdiv_(m, (v / bc2).sqrt_().add_(self.eps), value=-lr / bc1)
        self.p[sl], self.m[sl], self.v[sl] = p.to(self.p.dtype), m.to(self.p.dtype), v.to(self.p.dtype)
