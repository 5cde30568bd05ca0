"""Fused AdamW over flat bf16 parameter / gradient buffers (one kernel launch per step)."""
from __future__ import annotations

import torch

from . import _lib
from .ops import _p, _st


class FlatAdamW:
    """AdamW (torch.optim.AdamW semantics) for ONE flat bf16 parameter buffer whose .grad is a flat bf16 buffer; the
    model's parameters are views into those buffers (libra_b200.dist.FlatGradBuffer / bench.py).  States are bf16 like
    the parameters, matching what `model.to(bf16)` + torch AdamW gives in the reference recipe (train.py:31-32)."""

    def __init__(self, flat_param: torch.Tensor, flat_grad: torch.Tensor, lr=1e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0):
        assert flat_param.dtype == torch.bfloat16 and flat_grad.dtype == torch.bfloat16
        assert flat_param.numel() == flat_grad.numel()
        self.p, self.g = flat_param, flat_grad
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(flat_param)
        self.v = torch.zeros_like(flat_param)
        self.t = 0
        n = flat_param.numel()
        self.n_main = n - n % 8

    def step(self):
        self.t += 1
        _lib.call("lb_adamw_bf16", _p(self.p), _p(self.g), _p(self.m), _p(self.v), self.n_main, float(self.lr),
                  float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay), self.t, _st())
        if self.n_main != self.p.numel():          # < 8 trailing elements
            sl = slice(self.n_main, None)
            p, g = self.p[sl].float(), self.g[sl].float()
            m = self.m[sl].float().mul_(self.betas[0]).add_(g, alpha=1 - self.betas[0])
            v = self.v[sl].float().mul_(self.betas[1]).addcmul_(g, g, value=1 - self.betas[1])
            bc1, bc2 = 1 - self.betas[0] ** self.t, 1 - self.betas[1] ** self.t
            p.mul_(1 - self.lr * self.weight_decay).addcdiv_(m, (v / bc2).sqrt_().add_(self.eps), value=-self.lr / bc1)
            self.p[sl], self.m[sl], self.v[sl] = p.to(self.p.dtype), m.to(self.p.dtype), v.to(self.p.dtype)
