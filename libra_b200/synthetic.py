"""Synthetic workloads of SURVEY.md section 8(d): Libra batches with the reference's token layout
(<s> + image blocks of BOI/576 codes/EOI + text) and the bench weight randomisation."""
from __future__ import annotations

from typing import Dict, Optional

import torch

IMG = 578


def libra_batch(batch: int, seqlen: int, images_per_sample: int, vocab: int = 32000, signal: int = 2048, seed: int = 1234,
                device="cpu", signal_dtype=torch.bfloat16, pin: bool = False) -> Dict[str, torch.Tensor]:
    """input_ids [2,B,T], attention_mask [B,T] (all ones: padding-free throughput layout), vision_indices [B,T],
    contiguous_signal [B,T,signal], labels [2,B,T] following LibraTrainWrapper.get_labels
    (modeling_libra.py:1397-1411: -100 on BOS, BOI and the first text token after each image)."""
    g = torch.Generator().manual_seed(seed)
    T = seqlen
    assert 1 + images_per_sample * IMG <= T, "sequence too short for the requested images"
    ids = torch.randint(3, vocab, (batch, T), generator=g)
    ids[:, 0] = 1
    input_ids = ids[None].repeat(2, 1, 1)
    vi = torch.full((batch, T), IMG, dtype=torch.long)
    labels_mask = torch.zeros(batch, T, dtype=torch.bool)
    # images back to back right after <s> (cfg 3 / cfg 5 layout)
    pos = 1
    for _ in range(images_per_sample):
        input_ids[:, :, pos] = vocab + 512
        input_ids[:, :, pos + IMG - 1] = vocab + 513
        input_ids[:, :, pos + 1:pos + IMG - 1] = torch.randint(0, 512, (2, batch, IMG - 2), generator=g) + vocab
        vi[:, pos:pos + IMG] = torch.arange(IMG)
        if pos + IMG < T:
            labels_mask[:, pos + IMG] = True
        pos += IMG
    sig = torch.randn(batch, T, signal, generator=g).to(signal_dtype)
    sig[(vi >= IMG) | (vi == 0) | (vi == IMG - 1)] = 0
    labels = input_ids.clone()
    labels[labels == vocab + 512] = -100
    labels[labels == 1] = -100
    labels[:, labels_mask] = -100
    out = dict(input_ids=input_ids, attention_mask=torch.ones(batch, T, dtype=torch.long), vision_indices=vi,
               contiguous_signal=sig, labels=labels)
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    if device != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out


@torch.no_grad()
def randomize_for_bench(model, seed: int = 0, std: float = 0.02):
    """HF-style N(0, std) init is what the constructor did; additionally randomise every LibraLinear.weight_B (the
    bridge B matrices are zero-initialised, modeling_libra.py:506-507) and every norm weight, so that routing, low-rank
    and bridge paths are all numerically exercised."""
    g = torch.Generator(device=next(model.parameters()).device).manual_seed(seed)
    for n, p in model.named_parameters():
        if n.endswith("weight_B"):
            p.copy_(torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32) * std)
        elif "norm" in n and p.ndim == 1:
            p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32))


def decoder_flops(cfg, batch: int, seqlen: int, n_vis_per_sample: int, bridge_rank: int = 8) -> Dict[str, float]:
    """Model (algorithmic) forward FLOPs of one batch, SURVEY.md section 8(d)."""
    H, I, L, V, Vv, S = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.vocab_size, cfg.vision_vocab_size, cfg.contiguous_signal_size
    r = cfg.vision_down_ratio
    n_v = batch * n_vis_per_sample
    n_l = batch * seqlen - n_v
    bridge = 2 * 2 * 2 * H * bridge_rank                       # k and v bridges, A then B
    lang = 2 * (4 * H * H + 3 * H * I) + bridge
    lowrank = lambda i, o: 2 * (i * (o // r) + (o // r) * o)
    vis = 4 * lowrank(H, H) + 2 * lowrank(H, I) + lowrank(I, H) + bridge
    gemm = L * (n_l * lang + n_v * vis) + n_l * 2 * H * V + n_v * 2 * 2 * H * Vv + n_v * 2 * (H + S) * H
    attn = L * batch * 4 * seqlen * seqlen * H / 2            # causal-algorithmic QK^T + PV
    return dict(gemm=float(gemm), attn=float(attn), total=float(gemm + attn), attn_per_layer_fwd=float(attn / L))
