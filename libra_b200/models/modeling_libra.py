"""Host-side mirror of the reference's decoder interface (libra/models/libra/modeling_libra.py), running on the
libra_b200 CUDA kernels.

Same class names, constructor arguments, forward signature, output object and **state-dict keys** as the reference
(LibraLinear.weight_A/weight_B, self_attn.vision_k_bridge_on_language, model.vision_embed_tokens.{0,1}, ...), so
`from_pretrained` of a reference checkpoint, the reference's train.py and the HF Trainer work against it.

What differs is the execution plan: the [B,T,C] activations are permuted once per batch into
"sorted rows" (language tokens first, vision tokens second) and stay that way through all layers, which turns
the reference's ~52 boolean gather/scatter ops per layer (cal_language_vision, :111-147) into two contiguous row
ranges; attention runs on one fused tcgen05 kernel (libra_b200.functional.bridge_attention).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import torch
import torch.utils.checkpoint
from torch import nn
from transformers.modeling_outputs import CausalLMOutputWithPast
from transformers.modeling_utils import PreTrainedModel

from .. import _lib
from .. import ops
from .. import functional as LF
from .. import schedule
from ..registry import registry
from .configuration_libra import LibraConfig

BF16 = torch.bfloat16


@dataclass
class LibraCausalLMOutputWithPast(CausalLMOutputWithPast):
    """Same fields as the reference output class (modeling_libra.py:98-109)."""
    loss: Optional[torch.FloatTensor] = None
    logits: torch.FloatTensor = None
    past_key_values: Optional[Tuple[Tuple[torch.FloatTensor]]] = None
    hidden_states: Optional[Tuple[torch.FloatTensor]] = None
    attentions: Optional[Tuple[torch.FloatTensor]] = None
    past_hidden_states: Optional[torch.FloatTensor] = None
    past_vision_flag: Optional[torch.BoolTensor] = None


class LlamaRMSNorm(nn.Module):
    """Parameter holder (libra/models/llama/modeling_llama.py:118-132); the maths runs in lb_rmsnorm_*."""

    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):
        shp = x.shape
        return LF.rmsnorm(x.reshape(-1, shp[-1]).contiguous(), self.weight, None, None, self.variance_epsilon).view(shp)


class LibraLinear(nn.Module):
    """Low-rank pair y = (x A^T) B^T (modeling_libra.py:150-204): same parameters, shapes and initialisation."""

    def __init__(self, in_features: int, out_features: int, bias: bool = False, down_ratio=4, rank=None):
        super().__init__()
        assert in_features % down_ratio == 0
        assert bias is False, "Not checked yet."
        self.in_features, self.out_features, self.down_ratio, self.rank = in_features, out_features, down_ratio, rank
        mid = rank if rank is not None else out_features // down_ratio
        self.weight_A = nn.Parameter(torch.empty(mid, in_features))
        self.weight_B = nn.Parameter(torch.empty(out_features, mid))
        self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight_A, a=math.sqrt(5))
        if self.rank is not None:
            nn.init.constant_(self.weight_B, 0.0)
        else:
            nn.init.kaiming_uniform_(self.weight_B, a=math.sqrt(5))

    def forward(self, x):
        return LF.linear(LF.linear(x, self.weight_A), self.weight_B)


class _RotaryBuffers(nn.Module):
    """Keeps the `rotary_emb.inv_freq` state-dict key of the reference (modeling_llama.py:135-139)."""

    def __init__(self, dim, base=10000.0):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim)))


class LibraAttention(nn.Module):
    def __init__(self, config: LibraConfig):
        super().__init__()
        H = config.hidden_size
        self.hidden_size, self.num_heads = H, config.num_attention_heads
        self.head_dim = H // self.num_heads
        self.q_proj = nn.Linear(H, H, bias=False)
        self.k_proj = nn.Linear(H, H, bias=False)
        self.v_proj = nn.Linear(H, H, bias=False)
        self.o_proj = nn.Linear(H, H, bias=False)
        self.rotary_emb = _RotaryBuffers(self.head_dim)
        r = config.vision_down_ratio
        self.vision_q_proj = LibraLinear(H, H, down_ratio=r)
        self.vision_k_proj = LibraLinear(H, H, down_ratio=r)
        self.vision_v_proj = LibraLinear(H, H, down_ratio=r)
        self.vision_o_proj = LibraLinear(H, H, down_ratio=r)
        self.use_bridge = config.use_bridge
        if self.use_bridge:
            self.vision_v_bridge_on_language = LibraLinear(H, H, rank=config.bridge_rank)
            self.vision_v_bridge_on_vision = LibraLinear(H, H, rank=config.bridge_rank)
            self.vision_k_bridge_on_language = LibraLinear(H, H, rank=config.bridge_rank)
            self.vision_k_bridge_on_vision = LibraLinear(H, H, rank=config.bridge_rank)


class LibraMLP(nn.Module):
    def __init__(self, config: LibraConfig):
        super().__init__()
        H, I, r = config.hidden_size, config.intermediate_size, config.vision_down_ratio
        self.gate_proj = nn.Linear(H, I, bias=False)
        self.down_proj = nn.Linear(I, H, bias=False)
        self.up_proj = nn.Linear(H, I, bias=False)
        self.vision_gate_proj = LibraLinear(H, I, down_ratio=r)
        self.vision_down_proj = LibraLinear(I, H, down_ratio=r)
        self.vision_up_proj = LibraLinear(H, I, down_ratio=r)


class LibraDecoderLayer(nn.Module):
    """LibraDecoderLayer (modeling_libra.py:416-491) on sorted rows."""

    def __init__(self, config: LibraConfig):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = LibraAttention(config)
        self.mlp = LibraMLP(config)
        self.input_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.vision_input_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.vision_post_attention_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)

    def forward(self, h: torch.Tensor, meta: LF.AttnMeta) -> torch.Tensor:
        """h: [N, C] bf16, sorted rows."""
        rt = meta.routing
        nl, flag = rt.n_lang, rt.flag_sorted
        a, m = self.self_attn, self.mlp
        eps = self.input_layernorm.variance_epsilon
        # (residual branch, normed) behind one autograd node => the residual gradient is folded into the norm backward
        hr, n1 = LF.residual_rmsnorm(h, self.input_layernorm.weight, self.vision_input_layernorm.weight, flag, eps)
        q, k, v, tk, tv = LF.routed_fanout(
            n1, nl, ("lin", "lin", "lin", "down", "down"),
            a.q_proj.weight, a.vision_q_proj.weight_A, a.vision_q_proj.weight_B,
            a.k_proj.weight, a.vision_k_proj.weight_A, a.vision_k_proj.weight_B,
            a.v_proj.weight, a.vision_v_proj.weight_A, a.vision_v_proj.weight_B,
            a.vision_k_bridge_on_language.weight_A, a.vision_k_bridge_on_vision.weight_A,
            a.vision_v_bridge_on_language.weight_A, a.vision_v_bridge_on_vision.weight_A)
        o = LF.bridge_attention(q, k, v, tk, tv, a.vision_k_bridge_on_language.weight_B, a.vision_k_bridge_on_vision.weight_B,
                                a.vision_v_bridge_on_language.weight_B, a.vision_v_bridge_on_vision.weight_B, meta)
        # residual add fused into the output-projection GEMMs (beta = 1)
        h = LF.routed_linear(o, nl, a.o_proj.weight, a.vision_o_proj.weight_A, a.vision_o_proj.weight_B, residual=hr)
        hr, n2 = LF.residual_rmsnorm(h, self.post_attention_layernorm.weight, self.vision_post_attention_layernorm.weight, flag, eps)
        # gate | up with the SwiGLU product taken in the GEMM epilogue (language rows) / right behind the chains (vision rows)
        act = LF.routed_gate_up(n2, nl, m.gate_proj.weight, m.vision_gate_proj.weight_A, m.vision_gate_proj.weight_B,
                                m.up_proj.weight, m.vision_up_proj.weight_A, m.vision_up_proj.weight_B)
        return LF.routed_linear(act, nl, m.down_proj.weight, m.vision_down_proj.weight_A, m.vision_down_proj.weight_B, residual=hr)


class LibraPreTrainedModel(PreTrainedModel):
    config_class = LibraConfig
    base_model_prefix = "model"
    supports_gradient_checkpointing = True
    _no_split_modules = ["LibraDecoderLayer"]
    _skip_keys_device_placement = "past_key_values"

    def _init_weights(self, module):
        """modeling_libra.py:502-519."""
        std = self.config.initializer_range
        if isinstance(module, LibraLinear):
            module.weight_A.data.normal_(mean=0.0, std=std)
            if module.rank is not None or self.config.addition_mode:
                module.weight_B.data.zero_()
            else:
                module.weight_B.data.normal_(mean=0.0, std=std)
        elif isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()

    def _set_gradient_checkpointing(self, module=None, value=False, **kw):
        for m in self.modules():
            if isinstance(m, LibraModel):
                m.gradient_checkpointing = value

    def gradient_checkpointing_enable(self, gradient_checkpointing_kwargs=None):
        self._set_gradient_checkpointing(value=True)

    def gradient_checkpointing_disable(self):
        self._set_gradient_checkpointing(value=False)


class LibraModel(LibraPreTrainedModel):
    def __init__(self, config: LibraConfig):
        super().__init__(config)
        self.padding_idx = config.pad_token_id
        self.vocab_size = config.vocab_size
        H = config.hidden_size
        self.embed_tokens = nn.Embedding(config.vocab_size, H, self.padding_idx)
        self.layers = nn.ModuleList([LibraDecoderLayer(config) for _ in range(config.num_hidden_layers)])
        self.norm = LlamaRMSNorm(H, eps=config.rms_norm_eps)
        self.vision_vocab_size = config.vision_vocab_size
        self.vision_codebook_num = config.vision_codebook_num
        assert H % config.vision_codebook_num == 0
        self.vision_embed_tokens = nn.ModuleList(
            [nn.Embedding(config.vision_vocab_size, H // config.vision_codebook_num) for _ in range(config.vision_codebook_num)])
        self.vision_norm = LlamaRMSNorm(H, eps=config.rms_norm_eps)
        self.contiguous_signal_size = config.contiguous_signal_size
        self.vision_contiguous_signal_processor = nn.Linear(config.contiguous_signal_size + H, H, bias=False)
        self.vision_signal_norm = LlamaRMSNorm(config.contiguous_signal_size + H, eps=config.rms_norm_eps)
        self.max_vision_token_length = config.max_vision_token_length
        self.image_feature_resolution = config.image_feature_resolution
        assert self.image_feature_resolution ** 2 + 2 == self.max_vision_token_length
        self.gradient_checkpointing = False
        self._meta_cache = {}
        self._rope_cache = None
        self.post_init()

    def get_input_embeddings(self):
        return self.embed_tokens

    def set_input_embeddings(self, value):
        self.embed_tokens = value

    # -------------------------------------------------------------- per-batch metadata
    def _rope_tables(self, n_pos: int, device):
        c = self._rope_cache
        if c is None or c[0].shape[0] < n_pos or c[0].device != device:
            D = self.config.hidden_size // self.config.num_attention_heads
            inv_freq = 1.0 / (10000.0 ** (torch.arange(0, D, 2, device=device).float() / D))
            t = torch.arange(max(n_pos, self.config.max_position_embeddings), device=device, dtype=torch.float32)
            fr = torch.outer(t, inv_freq)
            # the reference rounds cos/sin to the activation dtype (bf16) before the rotation (modeling_llama.py:161-164)
            self._rope_cache = c = (fr.cos().to(BF16).float().contiguous(), fr.sin().to(BF16).float().contiguous())
        return c

    def build_meta(self, vision_flag: torch.Tensor, attention_mask: Optional[torch.Tensor],
                   position_ids: Optional[torch.Tensor], flag_cpu: Optional[torch.Tensor] = None,
                   am_cpu: Optional[torch.Tensor] = None) -> LF.AttnMeta:
        """Routing permutation, attention work lists, key ranges: built on the host once per distinct batch layout and
        cached.  The cache key is the layout itself; when the caller already holds it on the host (`flag_cpu`, `am_cpu`:
        LibraTokenizer and bench.py attach them to the device tensors, see schedule.attach_host_layout) nothing is copied
        back from the device, otherwise one device->host copy of the [B,T] flag (and mask) per call."""
        B, T = vision_flag.shape
        dev = vision_flag.device
        if flag_cpu is None:
            flag_cpu = vision_flag.detach().to("cpu")
        if attention_mask is None:
            am_cpu = None
        elif am_cpu is None:
            am_cpu = attention_mask.detach().to("cpu")
        am_cpu = None if am_cpu is None else am_cpu.to(torch.bool)
        key = (B, T, flag_cpu.numpy().tobytes(), None if am_cpu is None else am_cpu.numpy().tobytes())
        hit = self._meta_cache.get(key)
        if hit is None:
            kv_start = kv_end = None
            if am_cpu is not None and not bool(am_cpu.all()):
                kv_start, kv_end = [], []
                for b in range(B):
                    nz = torch.nonzero(am_cpu[b]).flatten()
                    s, e = (int(nz[0]), int(nz[-1]) + 1) if nz.numel() else (0, 0)
                    if nz.numel() != e - s:
                        raise NotImplementedError("attention_mask must be one contiguous run of ones per sample "
                                                  "(left or right padding)")
                    kv_start.append(s)
                    kv_end.append(e)
            rt_cpu = schedule.build_routing(flag_cpu)
            rt = schedule.Routing(rt_cpu.n_tokens, rt_cpu.n_lang, rt_cpu.n_vis, rt_cpu.perm.to(dev), rt_cpu.inv.to(dev),
                                  rt_cpu.flag_sorted.to(dev), rt_cpu.flag_orig.to(dev))
            work = schedule.build_attn_work(flag_cpu, B, T, True, dev, kv_start, kv_end)
            hit = (rt, work)
            if len(self._meta_cache) > 16:
                self._meta_cache.clear()
            self._meta_cache[key] = hit
        rt, work = hit
        if position_ids is None:
            pos = torch.arange(T, device=dev, dtype=torch.int32).repeat(B)
            n_pos = T
        else:
            pos = position_ids.to(dev).reshape(-1, T).expand(B, T).reshape(-1).to(torch.int32).contiguous()
            n_pos = int(pos.max().item()) + 1
        cos, sin = self._rope_tables(n_pos, dev)
        H = self.config.num_attention_heads
        return LF.AttnMeta(rt, work, pos, cos, sin, B, T, H, self.config.hidden_size // H)

    def build_decode_meta(self, vision_flag: torch.Tensor, attention_mask: Optional[torch.Tensor],
                          position_ids: Optional[torch.Tensor], cache) -> LF.AttnMeta:
        """Metadata of a one-token step (N1): B rows, one per sample (language rows first), the visible key range of every
        sample from its attention mask over past + new positions (modeling_libra.py:761-763 with a past)."""
        B = vision_flag.shape[0]
        dev = vision_flag.device
        Tk = cache.length + 1
        flag_cpu = vision_flag.detach().to("cpu")
        key = ("decode", B, flag_cpu.numpy().tobytes())
        rt = self._meta_cache.get(key)
        if rt is None:                                   # one entry per modality pattern of the B new tokens
            rt_cpu = schedule.build_routing(flag_cpu)
            rt = schedule.Routing(rt_cpu.n_tokens, rt_cpu.n_lang, rt_cpu.n_vis, rt_cpu.perm.to(dev), rt_cpu.inv.to(dev),
                                  rt_cpu.flag_sorted.to(dev), rt_cpu.flag_orig.to(dev))
            if len(self._meta_cache) > 64:
                self._meta_cache.clear()
            self._meta_cache[key] = rt
        kv_start = kv_end = None
        if attention_mask is not None:
            am = attention_mask.to(dev)
            if am.shape != (B, Tk):
                raise ValueError(f"attention_mask must cover past and new positions: expected {(B, Tk)}, got {tuple(am.shape)}")
            # first / last visible position per sample, computed on the device (no host round trip in the decode loop; that
            # the ones form one contiguous run was checked when the prompt went through build_meta)
            am8 = am.to(torch.int8)
            kv_start = am8.argmax(dim=1).to(torch.int32)
            kv_end = (Tk - am8.flip(1).argmax(dim=1)).to(torch.int32)
        if position_ids is None:
            pos = torch.full((B,), cache.length, device=dev, dtype=torch.int32)
        else:
            pos = position_ids.to(dev).reshape(-1).to(torch.int32).contiguous()
        cos, sin = self._rope_tables(cache.capacity + 1, dev)          # positions never exceed the cache length
        H = self.config.num_attention_heads
        kv_row = (torch.arange(B, device=dev, dtype=torch.int64) * cache.capacity + cache.len_dev).to(torch.int32)
        return LF.AttnMeta(rt, None, pos, cos, sin, B, 1, H, self.config.hidden_size // H, kv_cache=cache, decode=True,
                           dec_kv_start=kv_start, dec_kv_end=kv_end, dec_kv_row=kv_row)

    # -------------------------------------------------------------- embeddings (sorted rows)
    def embed_sorted(self, input_ids: torch.Tensor, meta: LF.AttnMeta, contiguous_signal: Optional[torch.Tensor]):
        """get_inputs_embeds_from_multicodebook (modeling_libra.py:625-661, :746-748) producing sorted rows."""
        rt = meta.routing
        nl = rt.n_lang
        perm = rt.perm.long()
        flat = input_ids.reshape(input_ids.shape[0], -1)
        h_lang = LF.EmbedLang.apply(flat[0][perm[:nl]].contiguous(), self.embed_tokens.weight, self.embed_tokens.padding_idx)
        if rt.n_vis == 0:
            return h_lang
        vis_rows = perm[nl:]
        ids0 = (flat[0][vis_rows] - self.vocab_size).contiguous()
        ids1 = (flat[1][vis_rows] - self.vocab_size).contiguous()
        sig = None
        if contiguous_signal is not None:
            sig = contiguous_signal.reshape(-1, contiguous_signal.shape[-1]).to(BF16).contiguous()
        cat = LF.EmbedVisionCat.apply(ids0, ids1, self.vision_embed_tokens[0].weight, self.vision_embed_tokens[1].weight, sig,
                                      rt.perm[nl:].contiguous(), self.contiguous_signal_size)
        cat = LF.rmsnorm(cat, self.vision_signal_norm.weight, None, None, self.vision_signal_norm.variance_epsilon)
        h_vis = LF.linear(cat, self.vision_contiguous_signal_processor.weight)
        return torch.cat([h_lang, h_vis], dim=0)

    def forward_sorted(self, input_ids, meta: LF.AttnMeta, contiguous_signal=None, collect_hidden=False):
        """Returns the final-normed hidden states in sorted rows [N, C] (and per-layer inputs if requested)."""
        h = self.embed_sorted(input_ids, meta, contiguous_signal)
        hiddens = [h] if collect_hidden else None
        hook = getattr(self, "layer_grad_ready_hook", None)
        for li, layer in enumerate(self.layers):
            if hook is not None and h.requires_grad:
                # fires when d(loss)/d(input of layer li) has been produced, i.e. after every gradient of layer li has been
                # enqueued: lets a data-parallel driver start reducing that layer's slice of the flat gradient buffer
                h.register_hook(lambda g, _li=li: hook(_li))
            meta.layer_idx = li
            if self.gradient_checkpointing and self.training and torch.is_grad_enabled():
                h = torch.utils.checkpoint.checkpoint(layer, h, meta, use_reentrant=False)
            else:
                h = layer(h, meta)
            if collect_hidden:
                hiddens.append(h)
        hn = LF.rmsnorm(h, self.norm.weight, self.vision_norm.weight, meta.routing.flag_sorted, self.norm.variance_epsilon)
        return hn, hiddens


class MultiLMHead(nn.Module):
    def __init__(self, head_num, input_dim, output_dim):
        super().__init__()
        self.heads = nn.ModuleList([nn.Linear(input_dim, output_dim, bias=False) for _ in range(head_num)])


class LibraForCausalLM(LibraPreTrainedModel):
    def __init__(self, config: LibraConfig):
        super().__init__(config)
        bad = config.unsupported_branches()
        if bad:
            raise NotImplementedError(
                "libra_b200 implements the reference's default (shipped) configuration; unsupported: " + ", ".join(bad))
        self.model = LibraModel(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.vision_codebook_num = config.vision_codebook_num
        self.vision_lm_head = MultiLMHead(config.vision_codebook_num, config.hidden_size, config.vision_vocab_size)
        self.max_vision_token_length = config.max_vision_token_length
        self.image_feature_resolution = config.image_feature_resolution
        # parameters/buffers the reference registers (kept for state-dict compatibility, :867-882)
        self.vision_hidden_placeholder = nn.Parameter(torch.empty(config.hidden_size))
        self.vision_hidden_placeholder.data.normal_(mean=0.0, std=config.initializer_range)
        self.register_buffer("naive_placeholder", torch.zeros(config.hidden_size))
        self.register_buffer("vision_logits_placeholder", torch.full([1, config.vision_vocab_size], -float("inf")))
        self.register_buffer("language_logits_placeholder", torch.full([1, config.vocab_size], -float("inf")))
        eoi = torch.full([1, 1, 1, config.vocab_size + config.vision_vocab_size], -float("inf"))
        eoi[:, :, :, config.newline_token_id] = float("inf")
        self.register_buffer("eoi_to_newline_logits_placeholder", eoi)
        self.post_init()

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def set_input_embeddings(self, value):
        self.model.embed_tokens = value

    def get_output_embeddings(self):
        return self.lm_head

    def get_decoder(self):
        return self.model

    # -------------------------------------------------------------- heads
    def _fused_loss(self, hn, meta, labels):
        """mean over codebooks of the shifted CE (modeling_libra.py:1159-1174) without materialising [Q,B,T,V+Vv]."""
        rt = meta.routing
        nl = rt.n_lang
        V, Vv, Q = self.config.vocab_size, self.config.vision_vocab_size, self.vision_codebook_num
        shift = torch.full_like(labels, -100)
        shift[:, :, :-1] = labels[:, :, 1:]
        ls = shift.reshape(Q, -1)[:, rt.perm.long()]                      # [Q, N] sorted rows
        counts = (ls != -100).sum(dim=1).to(torch.float32)                # per plane; 0 => 0/0 = nan like CrossEntropyLoss(mean)
        lab_l = ls[0, :nl]
        bad = ((lab_l >= V) & (lab_l != -100)).any()
        # one language-head pass serves every plane: the planes' text labels are copies of each other in the reference's
        # pipeline (get_labels clones the repeated text ids, :1397-1411); if a caller breaks that, fail loudly (nan), not silently
        planes_differ = (ls[1:, :nl] != ls[:1, :nl]).any() if Q > 1 else None
        S_l = LF.head_cross_entropy(hn[:nl], self.lm_head.weight, torch.where(lab_l >= V, -100, lab_l).contiguous())
        total = 0.0
        for c in range(Q):
            lab_v = ls[c, nl:]
            rel = lab_v - V
            oob = (lab_v != -100) & ((rel < 0) | (rel >= Vv))
            bad = bad | oob.any()
            S_v = LF.head_cross_entropy(hn[nl:], self.vision_lm_head.heads[c].weight,
                                        torch.where((lab_v == -100) | oob, -100, rel).contiguous()) if rt.n_vis else 0.0
            total = total + (S_l + S_v) / counts[c]
        loss = total / Q
        # a label outside its row's finite vocabulary block hits a -inf logit in the reference => loss = inf
        loss = loss + torch.where(bad, float("inf"), 0.0).to(loss.dtype)
        if planes_differ is not None:
            loss = loss + torch.where(planes_differ, float("nan"), 0.0).to(loss.dtype)
        return loss

    def _materialize_logits(self, hn, meta):
        """cal_vl_logits (modeling_libra.py:1018-1052): [Q,B,T,V+Vv] with -inf outside the row's block."""
        rt = meta.routing
        nl = rt.n_lang
        V, Vv, Q = self.config.vocab_size, self.config.vision_vocab_size, self.vision_codebook_num
        out = torch.full((Q, rt.n_tokens, V + Vv), float("-inf"), dtype=hn.dtype, device=hn.device)
        perm = rt.perm.long()
        ll = LF.linear(hn[:nl], self.lm_head.weight)
        for c in range(Q):
            out[c, perm[:nl], :V] = ll
            if rt.n_vis:
                out[c, perm[nl:], V:] = LF.linear(hn[nl:], self.vision_lm_head.heads[c].weight)
        return out.view(Q, meta.batch, meta.seqlen, V + Vv)

    def forward(
        self,
        input_ids: torch.LongTensor = None,
        attention_mask: Optional[torch.Tensor] = None,
        position_ids: Optional[torch.LongTensor] = None,
        past_key_values: Optional[List[torch.FloatTensor]] = None,
        inputs_embeds: Optional[torch.FloatTensor] = None,
        labels: Optional[torch.LongTensor] = None,
        use_cache: Optional[bool] = None,
        output_attentions: Optional[bool] = None,
        output_hidden_states: Optional[bool] = None,
        return_dict: Optional[bool] = None,
        vision_indices: Optional[torch.LongTensor] = None,
        contiguous_signal: Optional[torch.Tensor] = None,
        past_hidden_states: Optional[torch.Tensor] = None,
        past_vision_flag: Optional[torch.BoolTensor] = None,
        return_logits: Optional[bool] = None,
    ) -> Union[Tuple, LibraCausalLMOutputWithPast]:
        """Same arguments as the reference (modeling_libra.py:1069-1085).  `return_logits` (extra, optional):
        None = materialise the [Q,B,T,V+Vv] logits unless a training loss is being computed (HF Trainer only reads
        .loss); True/False force it."""
        _lib.require_device()
        if past_key_values is not None or use_cache:
            return self._forward_cached(input_ids, attention_mask, position_ids, past_key_values, vision_indices,
                                        contiguous_signal, labels, inputs_embeds, output_attentions, return_dict)
        if inputs_embeds is not None or output_attentions:
            raise NotImplementedError("inputs_embeds / output_attentions are not supported by the fused path")
        if input_ids is None or vision_indices is None:
            raise ValueError("input_ids [Q,B,T] and vision_indices [B,T] are required")
        if self.lm_head.weight.dtype != BF16:
            raise TypeError("libra_b200 computes in bf16: call model.to(torch.bfloat16) as train.py:31-32 does")
        assert len(input_ids) == self.vision_codebook_num
        vision_flag = vision_indices < self.max_vision_token_length
        flag_cpu, am_cpu = schedule.host_layout(vision_indices), schedule.host_layout(attention_mask)
        inconsistent = None
        if flag_cpu is None:
            # the reference's assertion (modeling_libra.py:708-711), eagerly: one host synchronisation
            if not torch.equal(vision_flag, input_ids[0] >= self.config.vocab_size):
                raise AssertionError("Inconsistent input_ids and vision_flag")
        else:
            # the layout came with a host copy: no device->host traffic in the steady state.  The same assertion is evaluated on
            # the device and, if violated, turns the loss / logits of this call into NaN (and raises at check_inputs()).
            inconsistent = (vision_flag != (input_ids[0] >= self.config.vocab_size)).any()
            self._inconsistent = inconsistent if getattr(self, "_inconsistent", None) is None else (self._inconsistent | inconsistent)
        meta = self.model.build_meta(vision_flag, attention_mask, position_ids, flag_cpu, am_cpu)
        hn, hiddens = self.model.forward_sorted(input_ids, meta, contiguous_signal, collect_hidden=bool(output_hidden_states))

        training_loss = labels is not None and torch.is_grad_enabled() and self.training
        want_logits = (not training_loss) if return_logits is None else return_logits
        loss = None
        if labels is not None:
            assert len(labels) == self.vision_codebook_num
            loss = self._fused_loss(hn, meta, labels.to(hn.device))
            if inconsistent is not None:
                loss = loss + torch.where(inconsistent, float("nan"), 0.0).to(loss.dtype)
        logits = self._materialize_logits(hn, meta) if want_logits else None
        if logits is not None and inconsistent is not None:
            logits = torch.where(inconsistent, torch.full_like(logits, float("nan")), logits)
        hs = None
        if hiddens is not None:
            inv = meta.routing.inv.long()
            C = hn.shape[-1]
            hs = tuple(t[inv].view(meta.batch, meta.seqlen, C) for t in hiddens[:-1]) + (hn[inv].view(meta.batch, meta.seqlen, C),)
        if return_dict is False:
            out = (logits,)
            return ((loss,) + out) if loss is not None else out
        return LibraCausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=hs,
                                           attentions=None, past_hidden_states=None, past_vision_flag=None)


    def check_inputs(self):
        """Raise the reference's "Inconsistent input_ids and vision_flag" assertion for calls that deferred it (one sync)."""
        bad = getattr(self, "_inconsistent", None)
        self._inconsistent = None
        if bad is not None and bool(bad):
            raise AssertionError("Inconsistent input_ids and vision_flag")

    # -------------------------------------------------------------- N1: use_cache=True (prefill and one-token steps)
    @torch.no_grad()
    def _forward_cached(self, input_ids, attention_mask, position_ids, past, vision_indices, contiguous_signal, labels,
                        inputs_embeds, output_attentions, return_dict):
        """forward(..., use_cache=True) of the reference (modeling_libra.py:1118-1145 over :680-831, :343-361).  Without a past:
        the whole prompt through the training kernels, every layer's key/value operands kept in a LibraKVCache.  With a
        past: ONE new token per sample (what prepare_inputs_for_generation :1190-1231 feeds) through lb_attn_decode."""
        from ..kv_cache import LibraKVCache
        if labels is not None or inputs_embeds is not None or output_attentions:
            raise NotImplementedError("use_cache=True is the inference path: no labels / inputs_embeds / output_attentions")
        if self.training:
            raise RuntimeError("use_cache=True needs model.eval() (the reference asserts `not self.training`, :1142)")
        if input_ids is None or vision_indices is None:
            raise ValueError("input_ids [Q,B,T] and vision_indices [B,T] are required")
        if self.lm_head.weight.dtype != BF16:
            raise TypeError("libra_b200 computes in bf16: call model.to(torch.bfloat16)")
        assert len(input_ids) == self.vision_codebook_num
        vision_flag = vision_indices < self.max_vision_token_length
        if not torch.equal(vision_flag, input_ids[0] >= self.config.vocab_size):
            raise AssertionError("Inconsistent input_ids and vision_flag")
        B, q_len = vision_flag.shape
        cfg = self.config
        if past is None:
            meta = self.model.build_meta(vision_flag, attention_mask, position_ids)
            cache = LibraKVCache(cfg.num_hidden_layers, B, q_len + 256, cfg.num_attention_heads,
                                 cfg.hidden_size // cfg.num_attention_heads, vision_flag.device)
            meta.kv_cache = cache
        else:
            if not isinstance(past, LibraKVCache):
                raise TypeError("past_key_values must be the LibraKVCache a previous use_cache=True call returned")
            if q_len != 1:
                raise NotImplementedError("with a past, one new token per sample (prepare_inputs_for_generation, :1190-1194)")
            cache = past
            cache.reserve(1)
            meta = self.model.build_decode_meta(vision_flag, attention_mask, position_ids, cache)
        try:
            # one-token steps are a chain of ~350 short launches: let each start while its predecessor drains (lb_set_pdl)
            with ops.pdl(past is not None):
                hn, _ = self.model.forward_sorted(input_ids, meta, contiguous_signal)
                cache.commit(vision_flag)
                logits = self._materialize_logits(hn, meta)
        finally:
            meta.kv_cache = None                       # the (cached) train-path metadata must not keep the cache alive
        if past is not None:
            # a row that just consumed </img> predicts nothing: "just append a newline" (:1142-1144)
            eoi = vision_indices[:, -1] == self.max_vision_token_length - 1
            logits = torch.where(eoi[None, :, None, None], self.eoi_to_newline_logits_placeholder.to(logits.dtype), logits)
        if return_dict is False:
            return (logits, cache)
        return LibraCausalLMOutputWithPast(loss=None, logits=logits, past_key_values=cache, hidden_states=None,
                                           attentions=None, past_hidden_states=None, past_vision_flag=None)

    @torch.no_grad()
    def generate(self, input_ids, attention_mask=None, vision_indices=None, contiguous_signal=None, max_new_tokens=32,
                 eos_token_id=None, pad_token_id=None, use_cache=True, do_sample=False, temperature=1.0, top_k=0, top_p=1.0,
                 repetition_penalty=1.0, logits_processor=None, logits_warper=None, generator=None, cuda_graph=True,
                 return_dict_in_generate=False, output_scores=False, num_beams=1, **unsupported):
        """Decoding with the KV cache, following the reference's generation plumbing: position_ids = cumsum(mask)-1
        (modeling_libra.py:1204-1205), the next token's vision index = previous + 1 inside an image, 578 after </img> or on
        text (:1273-1281), every codebook plane selected independently -- argmax (`greedy_search`,
        modeling_libra_utils.py:263-296) or, with do_sample=True, processors -> warpers -> softmax -> one multinomial draw
        per plane (`sample`, :538-564) -- and finished samples padded with pad_token_id (= eos_token_id when not given, as
        HF's generate does).  temperature / top_k / top_p / repetition_penalty are the transformers warpers
        (libra_b200/generation.py); `logits_processor` / `logits_warper` take any callables `(input_ids[B,T], scores[B,V])`
        (e.g. a transformers.LogitsProcessorList), applied per plane.  Returns the extended input_ids [Q,B,T+n], or a
        GenerateOutput with return_dict_in_generate=True.  Beam search is not part of this path.

        cuda_graph: once every sample is generating text (a language row can only predict text ids: the vision block of its
        logits is -inf), the one-token step -- ~350 short launches, launch-bound from Python -- is captured once in a CUDA
        graph with the cache addressed through its device-side length and the selected token fed back on the device, and
        replayed.  Caller-supplied processors, a caller-supplied generator and output_scores keep the eager loop (their
        state cannot be assumed capturable)."""
        from .. import generation as G
        if unsupported:
            raise TypeError(f"generate(): unsupported arguments {sorted(unsupported)} (this path covers greedy decoding and "
                            "sampling with the KV cache; see the docstring)")
        if num_beams != 1:
            raise NotImplementedError("beam search is outside this path (num_beams must be 1)")
        if vision_indices is None:
            raise ValueError("vision_indices [B,T] is required (578 on text positions), as in the reference's generate kwargs")
        if not use_cache:
            raise NotImplementedError("generate() decodes with the KV cache")
        policy = G.SelectionPolicy(do_sample=bool(do_sample), temperature=temperature, top_k=top_k, top_p=top_p,
                                   repetition_penalty=repetition_penalty, logits_processor=logits_processor,
                                   logits_warper=logits_warper, generator=generator)
        Q, B, T = input_ids.shape
        dev = input_ids.device
        eos_ids = None
        if eos_token_id is not None:
            eos_ids = torch.as_tensor([eos_token_id] if isinstance(eos_token_id, int) else list(eos_token_id), device=dev)
            if pad_token_id is None:
                pad_token_id = int(eos_ids[0])
        graph_ok = (cuda_graph and not policy.logits_processor and not policy.logits_warper and generator is None
                    and not output_scores)
        am = torch.ones(B, T, dtype=torch.long, device=dev) if attention_mask is None else attention_mask.to(dev).long()
        L = self.max_vision_token_length
        pos = am.cumsum(-1) - 1
        pos.masked_fill_(am == 0, 1)
        out = self.forward(input_ids=input_ids, attention_mask=am, position_ids=pos, vision_indices=vision_indices,
                           contiguous_signal=contiguous_signal, use_cache=True)
        done = torch.zeros(B, dtype=torch.bool, device=dev)
        vi_last = vision_indices[:, -1]
        n_done = 0
        all_scores = [] if output_scores else None

        def result(seq, cache):
            if not return_dict_in_generate:
                return seq
            return G.GenerateOutput(sequences=seq, scores=tuple(all_scores) if all_scores is not None else None,
                                    past_key_values=cache)

        while n_done < max_new_tokens:
            scores = G.process(policy, input_ids, out.logits[:, :, -1, :])
            if all_scores is not None:
                all_scores.append(scores)
            nxt = G.select(policy, scores)                                        # [Q, B]
            nxt, done = G.finish_(nxt, done, eos_ids, pad_token_id)
            input_ids = torch.cat([input_ids, nxt[:, :, None]], dim=2)
            n_done += 1
            vi_next = vi_last + 1
            vi_next = torch.where(vi_next >= L, torch.full_like(vi_next, L), vi_next)
            # a language token can open an image (<img> has vision index 0); keep flags consistent with the token
            is_vis_tok = nxt[0] >= self.config.vocab_size
            vi_next = torch.where(is_vis_tok & (vi_next >= L), torch.zeros_like(vi_next), vi_next)
            vi_next = torch.where(~is_vis_tok, torch.full_like(vi_next, L), vi_next)
            vi_last = vi_next
            am = torch.cat([am, am.new_ones(B, 1)], dim=1)
            if n_done >= max_new_tokens or (eos_ids is not None and bool(done.all())):
                break
            if graph_ok and max_new_tokens - n_done >= 4 and not bool(is_vis_tok.any()):
                # text from here on: replay the captured step for the remaining tokens
                toks = self._graph_decode(out.past_key_values, nxt, am, max_new_tokens - n_done, eos_ids, done,
                                          policy=policy, history=input_ids, pad_token_id=pad_token_id)
                return result(torch.cat([input_ids, toks], dim=2), out.past_key_values)
            p1 = (am.cumsum(-1) - 1)[:, -1:]
            out = self.forward(input_ids=nxt[:, :, None], attention_mask=am, position_ids=p1, vision_indices=vi_next[:, None],
                               past_key_values=out.past_key_values, use_cache=True)
        return result(input_ids, out.past_key_values)

    @torch.no_grad()
    def _graph_decode(self, cache, tokens, attention_mask, n_steps: int, eos_ids=None, done=None, policy=None, history=None,
                      pad_token_id=None):
        """Up to n_steps one-token steps on language tokens (greedy, or sampled under `policy`), starting from `tokens`
        [Q,B] (already appended to the sequence but not yet to the cache).  `history` [Q,B,T] is the sequence so far (the
        repetition penalty reads it; inside the graph it lives in a fixed-size buffer whose unused tail repeats each row's
        first token, which the penalty leaves unchanged).  The first step runs eagerly on a side stream (torch's capture warm-up, and a real
        step), the second is captured, the rest are replays.  With an EOS id the finished-sample bookkeeping runs inside the
        graph (a finished sample keeps emitting EOS, as in the eager loop) and the host looks at it every 16 replays.
        Returns the generated ids [Q,B,n] (n <= n_steps)."""
        from .. import generation as G
        from .. import schedule as _sch
        dev = tokens.device
        Q, B = tokens.shape
        cfg = self.config
        policy = policy or G.SelectionPolicy()
        hist = hist_pos = None
        if policy.repetition_penalty != 1.0:
            T0 = history.shape[2]
            hist = history[:, :, :1].repeat(1, 1, T0 + n_steps)
            hist[:, :, :T0] = history
            hist_pos = torch.full((1,), T0, dtype=torch.long, device=dev)
        cache.reserve(n_steps + 1)                                   # capacity (and every address) is fixed from here on
        H = cfg.num_attention_heads
        flag = torch.zeros(B, 1, dtype=torch.bool, device=dev)
        rt_cpu = _sch.build_routing(flag.cpu())
        rt = _sch.Routing(rt_cpu.n_tokens, rt_cpu.n_lang, rt_cpu.n_vis, rt_cpu.perm.to(dev), rt_cpu.inv.to(dev),
                          rt_cpu.flag_sorted.to(dev), rt_cpu.flag_orig.to(dev))
        am8 = attention_mask.to(torch.int8)
        kv_start = am8.argmax(dim=1).to(torch.int32)                 # left padding stays where it is
        kv_end = torch.zeros(B, dtype=torch.int32, device=dev)
        ids = tokens[:, :, None].clone()                             # static input of the step
        pos = ((attention_mask.cumsum(-1) - 1)[:, -1]).to(torch.int32).contiguous()
        outbuf = torch.zeros(Q, B, n_steps, dtype=torch.long, device=dev)
        step = torch.zeros(1, dtype=torch.long, device=dev)
        done = torch.zeros(B, dtype=torch.bool, device=dev) if done is None else done.clone()
        cos, sin = self.model._rope_tables(cache.capacity + 1, dev)
        kv_row = torch.zeros(B, dtype=torch.int32, device=dev)
        row0 = torch.arange(B, device=dev, dtype=torch.int64) * cache.capacity
        meta = LF.AttnMeta(rt, None, pos, cos, sin, B, 1, H, cfg.hidden_size // H, kv_cache=cache, decode=True,
                           dec_kv_start=kv_start, dec_kv_end=kv_end, dec_kv_row=kv_row)

        def body():
            kv_end.copy_((cache.len_dev + 1).to(torch.int32).expand(B))
            kv_row.copy_((row0 + cache.len_dev).to(torch.int32))
            with ops.pdl(True):                                      # captured as programmatic edges of the graph
                hn, _ = self.model.forward_sorted(ids, meta, None)
                cache.commit_device(flag)
                logits = self._materialize_logits(hn, meta)[:, :, -1]
            nxt = G.select(policy, G.process(policy, hist if hist is not None else ids, logits))
            G.finish_(nxt, done, eos_ids, pad_token_id)
            if hist is not None:
                hist.index_copy_(2, hist_pos, nxt[:, :, None])
                hist_pos.add_(1)
            outbuf.index_copy_(2, step, nxt[:, :, None])
            ids.copy_(nxt[:, :, None])
            pos.add_(1)
            step.add_(1)

        try:
            cur = torch.cuda.current_stream()
            s = torch.cuda.Stream()
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                body()                                               # step 1 (eager)
            cur.wait_stream(s)
            cache.length += 1
            n_run = 1
            if n_steps > 1:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    body()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n_first = n_run
                while n_run < n_steps:
                    if eos_ids is not None and n_run % 16 == 0 and bool(done.all()):
                        break
                    g.replay()
                    cache.length += 1
                    n_run += 1
                e1.record()
                self.last_graph_decode = (n_run - n_first, e0, e1)      # (replays, events around them): for benchmarks
        finally:
            meta.kv_cache = None
        return outbuf[:, :, :n_run]


@registry.register_model("libra_train_wrapper")
class LibraTrainWrapper(LibraPreTrainedModel):
    """LibraTrainWrapper (modeling_libra.py:1292-1437): tokenizer -> labels -> LibraForCausalLM, registered as
    "libra_train_wrapper" so `registry.get_model_class(arch).from_config(model_cfg)` (train.py:28-30) resolves to it.

    Two ways to build it:
      * the reference's: `LibraTrainWrapper(model_cfg)` with `model_cfg.pretrained` naming a checkpoint directory
        (config.json + weights + text tokenizer files + vision_tokenizer_config.yaml) and the optional `custom_kwargs`,
        `tokenizer_kwargs`, `model_kwargs`, `pretrained_weight` entries (:1294-1372); model_cfg may be an OmegaConf node, a
        dict or any object with `.get`;
      * injected: `LibraTrainWrapper(LibraConfig, module=..., tokenizer=..., model_kwargs=...)` -- the checkpoints are not
        part of the repository, synthetic runs and tests build the parts themselves.  The tokenizer is anything callable as
        `tokenizer(samples, return_tensors="pt", padding="longest", max_length=..., truncation=True)` returning `input_ids`,
        `attention_mask`, `vision_indices`, `coninous_signal` (libra_b200's LibraTokenizer or the reference's own)."""

    def __init__(self, config, module: Optional["LibraForCausalLM"] = None, tokenizer=None, model_kwargs=None):
        from transformers import PretrainedConfig
        get = (lambda k, d=None: config.get(k, d)) if (hasattr(config, "get") and not isinstance(config, PretrainedConfig)) else None
        if get is not None and get("pretrained") is not None:
            import json
            from pathlib import Path
            from .tokenization_libra import LibraTokenizer
            pretrained = str(get("pretrained"))
            super().__init__(LibraConfig(**json.load(open(Path(pretrained, "config.json"), "r"))))
            self.module = LibraForCausalLM.from_pretrained(pretrained, **dict(get("custom_kwargs", {}) or {}))
            self.tokenizer = LibraTokenizer(pretrained, **dict(get("tokenizer_kwargs", {}) or {}))
            tokenizer = self.tokenizer
            model_kwargs = dict(get("model_kwargs", {}) or {})
            weight = get("pretrained_weight", None)
            if weight is not None:                                             # :1312-1340
                sd = torch.load(weight, map_location="cpu")
                for prefix in ("model.model.", "module.model."):
                    if any(k.startswith(prefix) for k in sd):
                        cut = len(prefix) - len("model.")
                        sd = {k[cut:]: v for k, v in sd.items() if k.startswith(prefix[:cut])}
                        break
                missing, unexpected = self.module.load_state_dict(sd, strict=False)
                print("missing keys: ", missing)
                print("unexpected keys: ", unexpected)
        else:
            super().__init__(config)
            self.module = module if module is not None else LibraForCausalLM(config)
            self.tokenizer = tokenizer
        if tokenizer is not None:
            self.change_pad_token_to_eos(pad_token_id=tokenizer.text_tokenizer.pad_token_id,
                                         eos_token_id=tokenizer.text_tokenizer.eos_token_id)
        model_kwargs = model_kwargs or {}
        if model_kwargs.get("frozen_language", False):                      # :1342-1346
            for key, p in self.module.named_parameters():
                if "vision" not in key:
                    p.requires_grad = False
        for flag, pat in (("freeze_vision_value", "vision_v_proj"), ("freeze_text_embedding", ".embed_tokens"),
                          ("freeze_vision_embedding", ".vision_embed_tokens")):      # :1348-1364
            if model_kwargs.get(flag, False):
                for key, p in self.module.named_parameters():
                    if pat in key:
                        p.requires_grad = False
        if model_kwargs.get("debug", False):                                 # :1366-1369
            for key, p in self.module.named_parameters():
                if "vision_lm_head" not in key:
                    p.requires_grad = False

    @classmethod
    def get_model_from_config(cls, config):                                  # :1385-1391
        from .tokenization_libra import LibraTokenizer
        model = LibraForCausalLM.from_pretrained(config.pretrained, **dict(config.get("custom_kwargs", {}) or {}))
        return model, LibraTokenizer(config.pretrained, **dict(config.get("tokenizer_kwargs", {}) or {}))

    def get_optimizer_parameters(self):
        """Optional hook of LibraTrainer.create_optimizer (trainer.py:46-47, rewritten by :10-25): the default recipe's two
        groups -- weight decay on everything outside the norm modules that is not a bias."""
        from ..optim import decay_parameter_names
        decay = set(decay_parameter_names(self))
        named = [(n, p) for n, p in self.named_parameters() if p.requires_grad]
        return [{"params": [p for n, p in named if n in decay], "use_weight_decay": True},
                {"params": [p for n, p in named if n not in decay], "use_weight_decay": False}]

    @classmethod
    def from_config(cls, config, **kw):
        return cls(config, **kw)

    def change_pad_token_to_eos(self, pad_token_id=0, eos_token_id=2):       # :1390-1395
        w = self.module.get_input_embeddings().weight
        w.data[pad_token_id] = w.data[eos_token_id].clone()

    def get_labels(self, inputs, label_mask_position_map):                   # :1397-1411
        from .tokenization_libra import get_labels
        return get_labels(inputs["input_ids"], inputs["attention_mask"], self.tokenizer.image_tokenizer.boi_token_id,
                          self.tokenizer.text_tokenizer.bos_token_id, label_mask_position_map)

    def forward(self, samples, return_loss=None, **kwargs):                  # :1414-1433
        inputs = self.tokenizer(samples, return_tensors="pt", padding="longest",
                                max_length=self.tokenizer.text_tokenizer.model_max_length, truncation=True)
        labels = self.get_labels(inputs, samples["label_mask_position_map"])
        return self.module(input_ids=inputs["input_ids"], attention_mask=inputs["attention_mask"],
                           vision_indices=inputs["vision_indices"], contiguous_signal=inputs["coninous_signal"],
                           labels=labels, use_cache=False, **kwargs)
