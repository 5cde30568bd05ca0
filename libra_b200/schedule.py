"""Host-side scheduling metadata, computed once per batch (vision_flag is layer invariant,
libra/models/libra/modeling_libra.py:1118 -> :796,806):

  * the modality permutation (language rows first, vision rows second) that replaces the reference's
    per-call boolean gather/scatter (cal_language_vision, modeling_libra.py:111-147);
  * the attention work lists: (sample, 128-row q tile, variant) items, heaviest first.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

TILE = 128


@dataclass
class Routing:
    n_tokens: int
    n_lang: int
    n_vis: int
    perm: torch.Tensor          # [N] int32: sorted row r holds original token perm[r]
    inv: torch.Tensor           # [N] int32: original token bt lives in sorted row inv[bt]
    flag_sorted: torch.Tensor   # [N] uint8
    flag_orig: torch.Tensor     # [N] uint8


def build_routing(vision_flag: torch.Tensor) -> Routing:
    """vision_flag: [B,T] bool (any device).  One host sync per batch when it lives on the GPU."""
    dev = vision_flag.device
    f = vision_flag.reshape(-1)
    n = f.numel()
    # stable partition: language (False) first
    perm = torch.argsort(f.to(torch.int8), stable=True).to(torch.int32)
    inv = torch.empty_like(perm)
    inv[perm.long()] = torch.arange(n, dtype=torch.int32, device=dev)
    n_vis = int(f.sum().item())
    flag_sorted = f[perm.long()].to(torch.uint8)
    return Routing(n, n - n_vis, n_vis, perm.contiguous(), inv.contiguous(), flag_sorted.contiguous(),
                   f.to(torch.uint8).contiguous())


@dataclass
class AttnWork:
    work_q: torch.Tensor        # [n,4] int32 {b, q_tile, variant, 0} for forward / dQ
    work_kv: torch.Tensor       # [n,4] int32 {b, kv_tile, variant, first_q_tile} for dK/dV
    qtile_has: torch.Tensor     # [B,2,n_qtiles] uint8: q tile holds rows of modality v
    kv_start: Optional[torch.Tensor]
    kv_end: Optional[torch.Tensor]
    kv_cover: tuple = (False, False)   # per variant: every kv tile of every sample has a dK/dV work item
    q_tiles: Optional[list] = None           # kv tiles of every work_q item (host copy, for stream_plan)
    _plans: Optional[dict] = None
    kv_tiles: Optional[list] = None          # q tiles of every work_kv item (host copy, for stream_plan(which="kv"))

    def stream_plan(self, heads: int, n_cta: int, head_group: int = 8, overhead: float = 2.0, which: str = "q"):
        """Balanced split of the (work item, head) list over n_cta persistent CTAs (lb_attn_fwd_stream).
        Head groups are dealt in order (the K/V of one group stay L2-resident); inside a group the items go heaviest
        first to the least-loaded CTA, the load carried over from group to group.  weight = kv tiles + `overhead`
        (per-item switch cost in tile units).  Returns (plan_items [n_items], plan_off [n_cta+1] -- int32 on the work
        list's device --, n_cta, longest per-CTA list); cached per (heads, n_cta, head_group, which).
        which = "q": the work_q list (forward, dQ); "kv": the work_kv list (lb_attn_bwd_dkv_stream)."""
        import heapq
        if self._plans is None:
            self._plans = {}
        key = (heads, n_cta, head_group, which)
        tiles = self.q_tiles if which == "q" else self.kv_tiles
        if key not in self._plans:
            n_work = len(tiles)
            n_cta = max(1, min(n_cta, n_work * heads))
            loads = [(0.0, c) for c in range(n_cta)]
            heapq.heapify(loads)
            per_cta = [[] for _ in range(n_cta)]
            base = 0
            for g0 in range(0, heads, head_group):
                gl = min(head_group, heads - g0)
                for w in range(n_work):                       # work_q is sorted heaviest first
                    for hh in range(gl):
                        load, c = heapq.heappop(loads)
                        per_cta[c].append(base + w * gl + hh)
                        heapq.heappush(loads, (load + tiles[w] + overhead, c))
                base += head_group * n_work
            off = [0]
            for lst in per_cta:
                off.append(off[-1] + len(lst))
            items = torch.tensor([x for lst in per_cta for x in lst], dtype=torch.int32)
            dev = self.work_q.device
            self._plans[key] = (items.to(dev), torch.tensor(off, dtype=torch.int32).to(dev), n_cta,
                                max(len(lst) for lst in per_cta))
        return self._plans[key]


def build_attn_work(vision_flag_cpu: Optional[torch.Tensor], batch: int, seqlen: int, causal: bool, device,
                    kv_start=None, kv_end=None) -> AttnWork:
    """vision_flag_cpu: [B,T] bool on the host (None = a single variant, e.g. the ViT)."""
    nt = (seqlen + TILE - 1) // TILE
    has = torch.zeros(batch, 2, nt, dtype=torch.uint8)
    if vision_flag_cpu is None:
        has[:, 0, :] = 1
    else:
        f = vision_flag_cpu.reshape(batch, seqlen)
        pad = nt * TILE - seqlen
        fl = torch.nn.functional.pad(f.to(torch.uint8), (0, pad), value=2).view(batch, nt, TILE)
        has[:, 0] = (fl == 0).any(-1).to(torch.uint8)
        has[:, 1] = (fl == 1).any(-1).to(torch.uint8)
    ks = [0] * batch if kv_start is None else [int(x) for x in kv_start]
    ke = [seqlen] * batch if kv_end is None else [int(x) for x in kv_end]
    items_q, items_kv = [], []
    cover = [True, True]
    for b in range(batch):
        first_kv, last_kv = ks[b] // TILE, (ke[b] + TILE - 1) // TILE
        for v in range(2):
            for qt in range(nt):
                if not has[b, v, qt]:
                    continue
                n_kv = (min(last_kv, qt + 1) if causal else last_kv) - first_kv
                items_q.append((max(n_kv, 0), b, qt, v))
            for kt in range(first_kv, last_kv):
                first_q = kt if causal else 0
                n_q = int(has[b, v, first_q:].sum())
                if n_q > 0:
                    items_kv.append((n_q, b, kt, v, first_q))
                else:
                    cover[v] = False
            if first_kv > 0 or last_kv < nt:
                cover[v] = False
    items_q.sort(key=lambda t: -t[0])
    items_kv.sort(key=lambda t: -t[0])
    wq = torch.tensor([[b, qt, v, 0] for _, b, qt, v in items_q], dtype=torch.int32).reshape(-1, 4)
    wkv = torch.tensor([[b, kt, v, fq] for _, b, kt, v, fq in items_kv], dtype=torch.int32).reshape(-1, 4)
    to = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.int32).to(device)
    return AttnWork(wq.to(device), wkv.to(device), has.to(device), to(kv_start), to(kv_end), (cover[0], cover[1]),
                    [w for w, _, _, _ in items_q], None, [w for w, _, _, _, _ in items_kv])


# ----------------------------------------------------------------------------- host copies of the batch layout
def attach_host_layout(dev_tensor: torch.Tensor, host_tensor: torch.Tensor) -> torch.Tensor:
    """Remember the host copy of a layout tensor (vision flag of `vision_indices`, `attention_mask`) on the device tensor
    object that is handed to the model, so the model can key its per-layout metadata cache without a device->host copy.
    The producer of the batch (LibraTokenizer.forward: the text ids are tokenised on the host; bench.py: synthetic layouts)
    has these values on the host anyway.  For `vision_indices` the host tensor is the boolean vision flag [B,T]."""
    dev_tensor._lb_host_layout = host_tensor
    return dev_tensor


def host_layout(dev_tensor):
    if dev_tensor is None:
        return None
    h = getattr(dev_tensor, "_lb_host_layout", None)
    if h is not None and tuple(h.shape) != tuple(dev_tensor.shape):
        return None
    return h
