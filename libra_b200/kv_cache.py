"""KV cache of the decode path (N1): per layer the four attention operand tensors of the training kernels -- Kfl/Vfl (what
language queries see) and Kfv/Vfv (what vision queries see), post-RoPE, token-major [B, capacity, H*D] bf16 -- plus the
modality flag of every cached position.  The reference caches ([K_for_vision, K_for_language], V, V_bridge, vision_flag)
per layer (libra/models/libra/modeling_libra.py:354-361); `to_reference()` converts for inspection and the parity tests.

All samples of a batch advance together (generation batches are left padded), so one length serves the whole batch.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

BF16 = torch.bfloat16


class LibraKVCache:
    def __init__(self, n_layers: int, batch: int, capacity: int, heads: int, head_dim: int, device):
        self.n_layers, self.batch, self.capacity, self.heads, self.head_dim = n_layers, batch, capacity, heads, head_dim
        C = heads * head_dim
        mk = lambda: [torch.empty(batch, capacity, C, dtype=BF16, device=device) for _ in range(n_layers)]
        self.k_fl, self.v_fl, self.k_fv, self.v_fv = mk(), mk(), mk(), mk()
        self.flag = torch.zeros(batch, capacity, dtype=torch.bool, device=device)
        self.length = 0                  # positions filled in every layer
        self.len_dev = torch.zeros(1, dtype=torch.int64, device=device)      # the same number on the device: one-token
        # steps address the cache through it, so that a step can be captured in a CUDA graph and replayed
        self._pending = 0                # positions written by the layers of the step in flight

    # ---- HF-style accessors
    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self.length

    def __len__(self) -> int:
        return self.n_layers

    def __getitem__(self, i: int):
        return self.to_reference()[i]

    # ---- writes
    def reserve(self, n_new: int):
        """Make room for n_new more positions (doubling)."""
        need = self.length + n_new
        if need <= self.capacity:
            return
        cap = max(need, 2 * self.capacity)
        for lst in (self.k_fl, self.v_fl, self.k_fv, self.v_fv):
            for i, t in enumerate(lst):
                nt = torch.empty(self.batch, cap, t.shape[2], dtype=t.dtype, device=t.device)
                nt[:, :self.length] = t[:, :self.length]
                lst[i] = nt
        nf = torch.zeros(self.batch, cap, dtype=torch.bool, device=self.flag.device)
        nf[:, :self.length] = self.flag[:, :self.length]
        self.flag, self.capacity = nf, cap
        return True

    def append(self, layer: int, k_fv, k_fl, v_fv, v_fl, q_len: int):
        """Rows [B*q_len, C] in original token order (what lb_attn_prep_fwd writes) -> positions [length, length+q_len)."""
        B, s = self.batch, self.length
        for dst, src in ((self.k_fv, k_fv), (self.k_fl, k_fl), (self.v_fv, v_fv), (self.v_fl, v_fl)):
            if q_len == 1:
                dst[layer].index_copy_(1, self.len_dev, src.view(B, 1, -1))
            else:
                dst[layer][:, s:s + q_len].copy_(src.view(B, q_len, -1))
        self._pending = q_len

    def commit(self, flag_new: torch.Tensor):
        """All layers have appended the step's positions; flag_new [B, q_len] bool."""
        q = flag_new.shape[1]
        if q == 1:
            self.commit_device(flag_new)
        else:
            self.flag[:, self.length:self.length + q] = flag_new
            self.len_dev += q
        self.length += q
        self._pending = 0

    def commit_device(self, flag_new: torch.Tensor):
        """The device half of commit() for a one-token step (graph-capturable: no host state is touched)."""
        self.flag.index_copy_(1, self.len_dev, flag_new)
        self.len_dev += 1

    # ---- the reference's layout
    def to_reference(self) -> Tuple:
        """(([K_for_vision, K_for_language]), V, V_bridge, vision_flag) per layer, [B,H,T,hd] (modeling_libra.py:354-361)."""
        B, T, H, D = self.batch, self.length, self.heads, self.head_dim
        hd = lambda t: t[:, :T].view(B, T, H, D).transpose(1, 2)
        f = self.flag[:, :T]
        fk = f[:, None, :, None]
        out = []
        for i in range(self.n_layers):
            v_fv, v_fl = hd(self.v_fv[i]), hd(self.v_fl[i])
            v = torch.where(fk, v_fv, v_fl)                       # a key's own-modality value is the plain one
            vb = (torch.where(fk, v_fl, v_fv).float() - v.float()).to(v.dtype)
            out.append(([hd(self.k_fv[i]), hd(self.k_fl[i])], v, vb, f))
        return tuple(out)
