"""CLIP ViT vision tower on the libra_b200 kernels -- host mirror of libra/models/clip/modeling_clip.py
(CLIPVisionEmbeddings :170-228, CLIPAttention :262-363, CLIPMLP :366-378, CLIPEncoderLayer :381-428,
CLIPEncoder :600-700, CLIPVisionTransformer/CLIPVisionModel :859-972) with identical state-dict keys.

Kernels: im2col-free TMA-staged patch embedding (lb_patch_embed_fwd), LayerNorm fwd/bwd, non-causal tcgen05 flash
attention (head_dim 64) fwd/bwd; the dense projections (+bias, quick_gelu, residual in the epilogue) run on the grouped
tcgen05 GEMM (csrc/gemm_grouped.cu).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn
from transformers.configuration_utils import PretrainedConfig
from transformers.modeling_outputs import BaseModelOutputWithPooling
from transformers.modeling_utils import PreTrainedModel

from .. import _lib, ops, schedule
from .. import functional as LF

BF16 = torch.bfloat16


class CLIPVisionConfig(PretrainedConfig):
    """Fields/defaults of libra/models/clip/configuration_clip.py (vision part); ViT-L/14-336 when given those sizes."""
    model_type = "clip_vision_model"

    def __init__(self, hidden_size=768, intermediate_size=3072, projection_dim=512, num_hidden_layers=12,
                 num_attention_heads=12, num_channels=3, image_size=224, patch_size=32, hidden_act="quick_gelu",
                 layer_norm_eps=1e-5, attention_dropout=0.0, initializer_range=0.02, initializer_factor=1.0, **kwargs):
        super().__init__(**kwargs)
        self.hidden_size, self.intermediate_size, self.projection_dim = hidden_size, intermediate_size, projection_dim
        self.num_hidden_layers, self.num_attention_heads, self.num_channels = num_hidden_layers, num_attention_heads, num_channels
        self.patch_size, self.image_size = patch_size, image_size
        self.initializer_range, self.initializer_factor = initializer_range, initializer_factor
        self.attention_dropout, self.layer_norm_eps, self.hidden_act = attention_dropout, layer_norm_eps, hidden_act

    @classmethod
    def vit_l_14_336(cls):
        return cls(hidden_size=1024, intermediate_size=4096, projection_dim=768, num_hidden_layers=24, num_attention_heads=16,
                   image_size=336, patch_size=14)


class _PatchEmbed(torch.autograd.Function):
    """forward: lb_patch_embed_fwd.  backward (only exercised by the ViT-only fwd+bwd benchmark; Libra keeps the tower
    frozen, clip_encoder.py:27) through plain GEMMs on unfolded patches."""

    @staticmethod
    def forward(ctx, pixels, weight, class_emb, pos_emb):
        packed = ops.patch_embed_pack_weight(weight)
        ctx.save_for_backward(pixels)
        ctx.wshape = weight.shape
        return ops.patch_embed_fwd(pixels, packed, class_emb, pos_emb, patch=weight.shape[-1])

    @staticmethod
    def backward(ctx, dy):
        (pixels,) = ctx.saved_tensors
        C, _, P, _ = ctx.wshape
        B, _, S, _ = pixels.shape
        G = S // P
        K = 3 * P * P
        ld = (K + 7) // 8 * 8                                  # TMA row pitch: a multiple of 16 bytes (588 -> 592)
        patches = torch.zeros(B * G * G, ld, dtype=pixels.dtype, device=pixels.device)
        patches[:, :K] = pixels.reshape(B, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(B * G * G, K)
        dpatch = dy[:, 1:].reshape(B * G * G, C).contiguous()
        dWp = torch.empty(C, ld, dtype=pixels.dtype, device=pixels.device)
        ops.gemm_grouped([ops.gp(dpatch, patches[:, :K], dWp[:, :K], ta=True, tb=True)])     # dW = dpatch^T . patches
        dW = dWp[:, :K].reshape(C, 3, P, P)
        return None, dW, dy[:, 0].sum(0), dy.sum(0)


class CLIPVisionEmbeddings(nn.Module):
    def __init__(self, config: CLIPVisionConfig):
        super().__init__()
        self.embed_dim, self.image_size, self.patch_size = config.hidden_size, config.image_size, config.patch_size
        self.class_embedding = nn.Parameter(torch.randn(self.embed_dim))
        self.patch_embedding = nn.Conv2d(config.num_channels, self.embed_dim, kernel_size=self.patch_size, stride=self.patch_size, bias=False)
        self.num_patches = (self.image_size // self.patch_size) ** 2
        self.num_positions = self.num_patches + 1
        self.position_embedding = nn.Embedding(self.num_positions, self.embed_dim)
        self.register_buffer("position_ids", torch.arange(self.num_positions).expand((1, -1)))

    def forward(self, pixel_values):
        return _PatchEmbed.apply(pixel_values.to(BF16).contiguous(), self.patch_embedding.weight, self.class_embedding,
                                 self.position_embedding.weight)


class CLIPAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embed_dim, self.num_heads = config.hidden_size, config.num_attention_heads
        self.head_dim = self.embed_dim // self.num_heads
        self.scale = self.head_dim ** -0.5
        self.k_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.v_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.q_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.out_proj = nn.Linear(self.embed_dim, self.embed_dim)


class CLIPMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        if config.hidden_act != "quick_gelu":
            raise NotImplementedError("libra_b200 CLIP MLP implements quick_gelu (the ViT-L/14-336 activation)")
        self.fc1 = nn.Linear(config.hidden_size, config.intermediate_size)
        self.fc2 = nn.Linear(config.intermediate_size, config.hidden_size)


class CLIPEncoderLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.hidden_size
        self.self_attn = CLIPAttention(config)
        self.layer_norm1 = nn.LayerNorm(self.embed_dim, eps=config.layer_norm_eps)
        self.mlp = CLIPMLP(config)
        self.layer_norm2 = nn.LayerNorm(self.embed_dim, eps=config.layer_norm_eps)

    def forward(self, h, work, B, T):
        """h: [B*T, C] bf16."""
        a, m = self.self_attn, self.mlp
        x = LF.layernorm(h, self.layer_norm1.weight, self.layer_norm1.bias, self.layer_norm1.eps)
        # q / k / v (+bias): three problems of one grouped tcgen05 launch
        q, k, v = LF.linear_fanout(x, a.q_proj.weight, a.q_proj.bias, a.k_proj.weight, a.k_proj.bias, a.v_proj.weight, a.v_proj.bias)
        # the reference multiplies q by head_dim**-0.5 before q.k^T (:299); here the factor rides in the softmax scale
        o = LF.plain_attention(q, k, v, work, B, T, a.num_heads, a.head_dim, a.scale)
        h = LF.linear(o, a.out_proj.weight, a.out_proj.bias, residual=h)                # bias + residual in the epilogue
        x = LF.layernorm(h, self.layer_norm2.weight, self.layer_norm2.bias, self.layer_norm2.eps)
        f = LF.linear(x, m.fc1.weight, m.fc1.bias, act=1)                              # bias + quick_gelu in the epilogue
        return LF.linear(f, m.fc2.weight, m.fc2.bias, residual=h)


class CLIPEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleList([CLIPEncoderLayer(config) for _ in range(config.num_hidden_layers)])
        self.gradient_checkpointing = False


class CLIPVisionTransformer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = CLIPVisionEmbeddings(config)
        self.pre_layrnorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)    # (sic) reference attribute name
        self.encoder = CLIPEncoder(config)
        self.post_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self._work = {}

    def forward(self, pixel_values, output_hidden_states=False, last_layer: Optional[int] = None):
        _lib.require_device()
        emb = self.embeddings(pixel_values)
        B, T, C = emb.shape
        key = (B, T, emb.device)
        if key not in self._work:
            self._work[key] = schedule.build_attn_work(None, B, T, False, emb.device)
        work = self._work[key]
        h = LF.layernorm(emb.view(B * T, C), self.pre_layrnorm.weight, self.pre_layrnorm.bias, self.pre_layrnorm.eps)
        hs = [h.view(B, T, C)]
        layers = self.encoder.layers if last_layer is None else self.encoder.layers[:last_layer]
        for layer in layers:
            h = layer(h, work, B, T)
            hs.append(h.view(B, T, C))
        last = h.view(B, T, C)
        pooled = LF.layernorm(last[:, 0].contiguous(), self.post_layernorm.weight, self.post_layernorm.bias, self.post_layernorm.eps)
        return BaseModelOutputWithPooling(last_hidden_state=last, pooler_output=pooled,
                                          hidden_states=tuple(hs) if output_hidden_states else None, attentions=None)


class CLIPVisionModel(PreTrainedModel):
    config_class = CLIPVisionConfig
    main_input_name = "pixel_values"
    base_model_prefix = "clip"

    def __init__(self, config: CLIPVisionConfig):
        super().__init__(config)
        self.vision_model = CLIPVisionTransformer(config)
        self.post_init()

    def _init_weights(self, module):
        """CLIPPreTrainedModel._init_weights for the vision modules (modeling_clip.py:431-480)."""
        f = self.config.initializer_factor
        if isinstance(module, CLIPVisionEmbeddings):
            nn.init.normal_(module.class_embedding, mean=0.0, std=module.embed_dim ** -0.5 * f)
            nn.init.normal_(module.patch_embedding.weight, std=self.config.initializer_range * f)
            nn.init.normal_(module.position_embedding.weight, std=self.config.initializer_range * f)
        elif isinstance(module, CLIPAttention):
            in_std = (module.embed_dim ** -0.5) * ((2 * self.config.num_hidden_layers) ** -0.5) * f
            out_std = (module.embed_dim ** -0.5) * f
            for lin, s in ((module.q_proj, in_std), (module.k_proj, in_std), (module.v_proj, in_std), (module.out_proj, out_std)):
                nn.init.normal_(lin.weight, std=s)
                nn.init.zeros_(lin.bias)
        elif isinstance(module, CLIPMLP):
            in_std = (self.config.hidden_size ** -0.5) * ((2 * self.config.num_hidden_layers) ** -0.5) * f
            nn.init.normal_(module.fc1.weight, std=(2 * self.config.hidden_size) ** -0.5 * f)
            nn.init.normal_(module.fc2.weight, std=in_std)
            nn.init.zeros_(module.fc1.bias)
            nn.init.zeros_(module.fc2.bias)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    def get_input_embeddings(self):
        return self.vision_model.embeddings.patch_embedding

    def forward(self, pixel_values=None, output_attentions=None, output_hidden_states=None, return_dict=None, last_layer=None):
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised by the fused kernel")
        return self.vision_model(pixel_values, output_hidden_states=bool(output_hidden_states), last_layer=last_layer)
