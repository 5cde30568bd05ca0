"""Vision tokenizer encode + input tensor assembly on the GPU -- host mirror of
  * CLIPVisionTower (libra/models/libra/clip_encoder.py:31-69),
  * VQModel.encode (libra/models/libra/taming/models/vqgan.py:106-114) with quant_conv 1x1 and the lookup-free
    quantizer in eval mode (taming/modules/quantization/lookup_free_quantization.py:185-208),
  * ImageTokenizer.encode (libra/models/libra/image_tokenizer.py:75-95),
  * the tensor assembly half of LibraTokenizer.forward (libra/models/libra/tokenization_libra.py:250-316) and
    LibraTrainWrapper.get_labels (libra/models/libra/modeling_libra.py:1397-1411).
Text tokenisation (sentencepiece) and image decoding are out of scope: callers pass token ids.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from .. import _lib, ops
from .. import functional as LF
from .modeling_clip import CLIPVisionConfig, CLIPVisionModel

BF16 = torch.bfloat16


class LFQ(nn.Module):
    """Parameter holder with the reference's names (project_in/project_out when dim != codebook_dims)."""

    def __init__(self, dim: Optional[int], codebook_size: int = 512, num_codebooks: int = 2):
        super().__init__()
        self.codebook_dim = int(math.log2(codebook_size))
        assert 2 ** self.codebook_dim == codebook_size
        self.num_codebooks = num_codebooks
        codebook_dims = self.codebook_dim * num_codebooks
        self.dim = dim if dim is not None else codebook_dims
        self.has_projections = self.dim != codebook_dims
        self.project_in = nn.Linear(self.dim, codebook_dims) if self.has_projections else nn.Identity()
        self.project_out = nn.Linear(codebook_dims, self.dim) if self.has_projections else nn.Identity()
        self.register_buffer("mask", 2 ** torch.arange(self.codebook_dim - 1, -1, -1))


class VisionTokenizer(nn.Module):
    """ImageTokenizer + VQModel(encode side): frozen CLIP tower -> quant_conv -> LFQ indices -> ids with BOI/EOI.

    `select_layer`, `embed_dim` live in the un-shipped vision_tokenizer_config.yaml and are
    therefore constructor arguments here."""

    def __init__(self, clip_config: CLIPVisionConfig, select_layer: Sequence[int] = (-2, -6), embed_dim: int = 18,
                 codebook_size: int = 512, num_codebook: int = 2, token_offset: int = 32000):
        super().__init__()
        self.encoder = CLIPVisionModel(clip_config)
        self.select_layer = list(select_layer) if isinstance(select_layer, (list, tuple)) else [select_layer]
        self.quant_conv = nn.Conv2d(clip_config.hidden_size * len(self.select_layer), embed_dim, 1)
        self.quantize = LFQ(embed_dim, codebook_size, num_codebook)
        self.codebook_size, self.num_codebook = codebook_size, num_codebook
        self.offset = token_offset
        self.boi_token_id = token_offset + codebook_size
        self.eoi_token_id = token_offset + codebook_size + 1
        self.grid = clip_config.image_size // clip_config.patch_size
        self.max_vision_token_length = self.grid ** 2 + 2
        self.requires_grad_(False)
        self.eval()

    def train(self, mode=True):       # frozen + eval-locked like the reference (image_tokenizer.py:37-42)
        return super().train(False)

    @torch.no_grad()
    def encode(self, pixel_values: torch.Tensor) -> Dict[str, torch.Tensor]:
        _lib.require_device()
        n_layers = len(self.encoder.vision_model.encoder.layers)
        # hidden_states index -> number of layers to run; skip the layers nobody selects (clip_encoder.py:31-45)
        need = max((i if i >= 0 else n_layers + 1 + i) for i in self.select_layer)
        out = self.encoder(pixel_values, output_hidden_states=True, last_layer=need)
        hs = out.hidden_states
        full = lambda i: hs[i if i >= 0 else i + n_layers + 1]
        feat = torch.cat([full(i) for i in self.select_layer], dim=-1)[:, 1:].contiguous()      # [B, 576, C*len]
        B, N, Cin = feat.shape
        w = self.quant_conv.weight.view(self.quant_conv.out_channels, Cin)
        h = LF.linear(feat.view(B * N, Cin), w, self.quant_conv.bias).contiguous()            # 1x1 conv == per-token linear
        if self.quantize.has_projections:
            h = LF.linear(h, self.quantize.project_in.weight, self.quantize.project_in.bias)
        ids = ops.lfq_pack(h.contiguous(), B, N, self.num_codebook, self.quantize.codebook_dim, self.offset, self.boi_token_id,
                           self.eoi_token_id)
        return {"input_ids": ids, "image_size": [self.grid, self.grid],
                "attention_mask": torch.ones(ids.shape[1:], dtype=torch.long, device=ids.device), "encoder_feat": feat,
                "pre_quant": h}

    forward = encode


@torch.no_grad()
def assemble_inputs(text_ids: torch.Tensor, attention_mask: torch.Tensor, img_ph_token_id: int, image_ids: Optional[torch.Tensor],
                    encoder_feat: Optional[torch.Tensor], max_vision_token_length: int = 578,
                    contiguous_ignore: Optional[torch.Tensor] = None, truncation: bool = False,
                    max_length: Optional[int] = None, check: bool = False) -> Dict[str, torch.Tensor]:
    """tokenization_libra.py:250-316 on already-tokenised text: text_ids [B,T] hold `img_ph_token_id` at the 578
    placeholder positions of every image (batch-major image order).  Output keys as the reference, including the
    misspelt `coninous_signal`.

    The reference scatters through boolean masks (`ids[:, ph] = ...`: a `nonzero` and a device->host sync each, three per
    batch, between the vision tokenizer and the decoder).  Here the k-th placeholder position of the flattened batch reads
    element k of the flattened image tensors: k = cumsum(ph) - 1, gathers and `where`s only -- nothing leaves the device
    (SURVEY.md section 8f, N4).  check=True verifies that the placeholder count equals the number of image tokens (one
    sync; the reference's scatter raises on a mismatch)."""
    Q = 2 if image_ids is None else image_ids.shape[0]
    L = max_vision_token_length
    B, T = text_ids.shape
    if image_ids is None:
        ids = text_ids[None].repeat(Q, 1, 1)
        vi = torch.full(text_ids.shape, L, dtype=torch.long, device=text_ids.device)
        sig = None
    else:
        ph = text_ids == img_ph_token_id
        n_tok = image_ids.shape[1] * image_ids.shape[2]
        if check and int(ph.sum()) != n_tok:
            raise ValueError(f"{int(ph.sum())} placeholder positions for {n_tok} image tokens")
        rank = (ph.reshape(-1).cumsum(0) - 1).clamp_(0, max(n_tok - 1, 0)).view(B, T)        # k-th placeholder -> image token k
        ids = torch.where(ph[None], image_ids.flatten(1, 2)[:, rank.reshape(-1)].view(Q, B, T), text_ids[None])
        vi = torch.where(ph, rank % L, torch.full_like(rank, L))
        z = encoder_feat.new_zeros(encoder_feat.shape[0], 1, encoder_feat.shape[2])
        cont = torch.cat([z, encoder_feat, z], dim=1)                                          # zero rows at <img> and </img>
        if contiguous_ignore is not None:
            cont = torch.where(contiguous_ignore.view(-1, 1, 1).to(torch.bool), torch.zeros_like(cont), cont)
        sig = torch.where(ph[..., None], cont.flatten(0, 1)[rank.reshape(-1)].view(B, T, -1), torch.zeros((), dtype=cont.dtype, device=cont.device))
    if truncation and max_length is not None:
        ids, attention_mask, vi = ids[:, :, :max_length], attention_mask[:, :max_length], vi[:, :max_length]
        sig = None if sig is None else sig[:, :max_length]
    return {"input_ids": ids.contiguous(), "attention_mask": attention_mask.contiguous(), "vision_indices": vi.contiguous(),
            "coninous_signal": None if sig is None else sig.contiguous()}


@torch.no_grad()
def get_labels(input_ids: torch.Tensor, attention_mask: torch.Tensor, boi_token_id: int, bos_token_id: int,
               label_mask_position_map: Sequence[Sequence[Sequence[int]]]) -> torch.Tensor:
    """LibraTrainWrapper.get_labels (modeling_libra.py:1397-1411).  The span list is host data: it becomes ONE boolean mask
    built on the host and one `where` on the device, instead of a slice assignment (kernel launch) per span."""
    B, T = attention_mask.shape
    span = torch.zeros(B, T, dtype=torch.bool)
    for b, spans in enumerate(label_mask_position_map):
        for (s, e) in spans:
            span[b, s:e] = True
    drop = (attention_mask == 0) | span.to(input_ids.device, non_blocking=True)
    drop = drop[None] | (input_ids == boi_token_id) | (input_ids == bos_token_id)
    return torch.where(drop, torch.full_like(input_ids, -100), input_ids)
