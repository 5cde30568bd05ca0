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

    @property
    def device(self):
        return self.quant_conv.weight.device

    @property
    def dtype(self):
        return self.quant_conv.weight.dtype

    def __len__(self) -> int:           # image_tokenizer.py:59-60
        return self.codebook_size + 2

    def get_token_length(self, images=None) -> int:      # image_tokenizer.py:62-68 (fixed-size images)
        return self.max_vision_token_length

    @classmethod
    def from_config(cls, config, token_offset: int):
        """ImageTokenizer.from_config on the dict read from vision_tokenizer_config.yaml (image_tokenizer.py:16-48,
        taming/models/vqgan.py:44-75): keys params.{ddconfig.{encoder_name, select_layer}, embed_dim, codebook_size,
        num_codebook, ckpt_path}, max_vision_token_length.  The CLIP tower is loaded from `encoder_name`; `ckpt_path`
        (optional) supplies quant_conv / LFQ projection weights."""
        import os
        params = config["params"]
        dd = params["ddconfig"]
        enc_dir = dd["encoder_name"]
        clip_cfg = CLIPVisionConfig.from_pretrained(enc_dir)
        tok = cls(clip_cfg, select_layer=dd.get("select_layer", -2), embed_dim=params.get("embed_dim", 18),
                  codebook_size=params.get("codebook_size", 512), num_codebook=params.get("num_codebook", 2), token_offset=token_offset)
        if any(f.endswith((".safetensors", ".bin")) for f in os.listdir(enc_dir)):
            tok.encoder = CLIPVisionModel.from_pretrained(enc_dir)
        ck = params.get("ckpt_path")
        if ck and os.path.exists(ck):
            sd = torch.load(ck, map_location="cpu")
            sd = sd.get("state_dict", sd)
            own = tok.state_dict()
            tok.load_state_dict({k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}, strict=False)
        if config.get("max_vision_token_length") not in (None, tok.max_vision_token_length):
            raise ValueError("max_vision_token_length of the config does not match the CLIP grid")
        tok.requires_grad_(False)
        return tok.eval()

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

    # ------------------------------------------------------------------ N2: ids -> pixels (image_tokenizer.py:97-124)
    def attach_decoder(self, ddconfig: Dict, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        """Build the decode side (post_quant_conv + taming Decoder, models/vq_decoder.py) from the `ddconfig` of
        vision_tokenizer_config.yaml; `state_dict`: VQModel checkpoint keys (decoder.*, post_quant_conv.*, quantize.project_out.*)."""
        from .vq_decoder import VQDecoder
        dec = VQDecoder(ddconfig, embed_dim=self.quant_conv.out_channels, codebook_size=self.codebook_size,
                        num_codebook=self.num_codebook, token_offset=self.offset)
        if state_dict is not None:
            own = dec.state_dict()
            dec.load_state_dict({k: v for k, v in state_dict.items() if k in own}, strict=True)
        self.vq_decoder = dec.to(self.dtype).to(self.device)
        return self.vq_decoder

    @torch.no_grad()
    def decode(self, x):
        """ImageTokenizer.decode: token ids [Q,B,N] / [Q,N] (list or tensor, with or without <img> </img>) -> pixels."""
        if len(x) == 0 or len(x[0]) == 0:
            return x
        if not isinstance(x, torch.Tensor):
            x = torch.tensor(x, dtype=torch.long, device=self.device)
        if getattr(self, "vq_decoder", None) is None:
            raise RuntimeError("VisionTokenizer.decode: no decoder attached (attach_decoder(ddconfig, state_dict))")
        return self.vq_decoder.decode_ids(x.to(self.device))


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


# ----------------------------------------------------------------------------------------------------------------------
# LibraTokenizer (libra/models/libra/tokenization_libra.py:108-316): text tokenizer + vision tokenizer -> model inputs
# ----------------------------------------------------------------------------------------------------------------------
class SimpleTextTokenizer:
    """A dependency-free stand-in for the reference's LibraTextTokenizer (a sentencepiece LlamaTokenizer whose model file ships
    with the checkpoint, not with the repository): whitespace words hashed into the Llama id range, with the attributes
    and call signature LibraTokenizer / LibraTrainWrapper use (`__call__(texts, return_tensors="pt", return_length=True,
    padding=...)`, `add_tokens`, `convert_tokens_to_ids`, `batch_decode`, `vocab_size`, `bos/eos/unk/pad_token_id`,
    `padding_side`, `add_eos_token`, `model_max_length`).  For synthetic benches and tests; with a real checkpoint directory
    LibraTokenizer loads the Hugging Face tokenizer instead."""

    def __init__(self, vocab_size: int = 32000, model_max_length: int = 2048, padding_side: str = "right", add_eos_token: bool = False):
        self.vocab_size = vocab_size
        self.model_max_length = model_max_length
        self.padding_side = padding_side
        self.add_eos_token = add_eos_token
        self.unk_token, self.bos_token, self.eos_token = "<unk>", "<s>", "</s>"
        self.unk_token_id, self.bos_token_id, self.eos_token_id = 0, 1, 2
        self.pad_token = None
        self.added = {}
        self._words = {}

    def __len__(self):
        return self.vocab_size + len(self.added)

    @property
    def pad_token_id(self):
        return {None: None, self.unk_token: 0, self.bos_token: 1, self.eos_token: 2}.get(self.pad_token, None)

    def add_tokens(self, tok):
        if tok not in self.added:
            self.added[tok] = self.vocab_size + len(self.added)
        return 1

    def convert_tokens_to_ids(self, tok):
        return self.added.get(tok, self._word_id(tok))

    def _word_id(self, w):
        import zlib
        i = 3 + zlib.crc32(w.encode("utf-8")) % (self.vocab_size - 3)
        self._words.setdefault(i, w)
        return i

    def __call__(self, texts, return_tensors="pt", return_length=False, padding=False, **kw):
        from transformers import BatchEncoding
        if isinstance(texts, str):
            texts = [texts]
        rows = []
        for t in texts:
            ids = [self.bos_token_id] + [self.added.get(w, None) if w in self.added else self._word_id(w) for w in t.split()]
            if self.add_eos_token:
                ids.append(self.eos_token_id)
            rows.append(ids)
        n = max(len(r) for r in rows)
        pad = self.pad_token_id if self.pad_token_id is not None else 0
        ids = torch.full((len(rows), n), pad, dtype=torch.long)
        am = torch.zeros(len(rows), n, dtype=torch.long)
        for i, r in enumerate(rows):
            sl = slice(n - len(r), n) if self.padding_side == "left" else slice(0, len(r))
            ids[i, sl] = torch.tensor(r)
            am[i, sl] = 1
        out = {"input_ids": ids, "attention_mask": am}
        if return_length:
            out["length"] = torch.tensor([len(r) for r in rows])
        return BatchEncoding(out)

    def batch_decode(self, ids, skip_special_tokens=True, **kw):
        inv = {v: k for k, v in self.added.items()}
        outs = []
        for row in torch.as_tensor(ids).tolist():
            ws = []
            for i in row:
                if skip_special_tokens and i in (0, 1, 2):
                    continue
                ws.append(inv.get(i, self._words.get(i, f"<{i}>")))
            outs.append(" ".join(ws))
        return outs


def _cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default) if not hasattr(cfg, "get") else cfg.get(key, default)


class LibraTokenizer(torch.nn.Module):
    """tokenization_libra.py:108-316.  `LibraTokenizer(pretrained_model_path, vision_config_overwrite={}, **tok_kwargs)` reads
    the Hugging Face text tokenizer and `vision_tokenizer_config.yaml` from a checkpoint directory like the reference;
    alternatively the two tokenizers are injected (`text_tokenizer=`, `image_tokenizer=`) -- the checkpoints are not part of
    the repository.  forward(samples) returns the reference's keys, `coninous_signal` (sic) included.

    The vision side runs on this library's kernels (VisionTokenizer: CLIP tower, quant_conv, LFQ bit pack) and the tensor
    assembly is `assemble_inputs`: rank-of-placeholder gathers instead of the reference's three boolean-mask scatters, i.e.
    no `nonzero` and no device->host synchronisation between the tokenizer and the decoder."""

    def __init__(self, pretrained_model_path=None, vision_config_overwrite=None, text_tokenizer=None, image_tokenizer=None, **kwargs):
        super().__init__()
        self.raw_output = kwargs.pop("raw_output", False)
        if text_tokenizer is None:
            if pretrained_model_path is None:
                raise ValueError("LibraTokenizer needs a checkpoint directory or an injected text_tokenizer")
            text_tokenizer = self.init_text_tokenizer(pretrained_model_path, **kwargs)
        else:
            self._prepare_text_tokenizer(text_tokenizer)
        self.text_tokenizer = text_tokenizer
        self.image_tokenizer_offset = self.text_tokenizer.vocab_size
        if image_tokenizer is None:
            if pretrained_model_path is None:
                raise ValueError("LibraTokenizer needs a checkpoint directory or an injected image_tokenizer")
            image_tokenizer = self.init_image_tokenizer(pretrained_model_path, self.image_tokenizer_offset, vision_config_overwrite or {})
        self.image_tokenizer = image_tokenizer
        L = self.image_tokenizer.max_vision_token_length
        self.register_buffer("img_indices_ph", torch.arange(0, L, dtype=torch.long)[None, :])
        self.num_codebook = self.image_tokenizer.num_codebook

    @property
    def device(self):
        return self.image_tokenizer.device

    @property
    def dtype(self):
        return self.image_tokenizer.dtype

    @staticmethod
    def _prepare_text_tokenizer(tok):                                   # :141-151
        tok.add_tokens("<img_ph>")
        tok.add_tokens("<img_gen>")
        tok.img_ph_token_id = tok.convert_tokens_to_ids("<img_ph>")
        tok.img_gen_token_id = tok.convert_tokens_to_ids("<img_gen>")
        tok.pad_token = tok.unk_token
        return tok

    @classmethod
    def init_text_tokenizer(cls, pretrained_model_path, **kwargs):
        from transformers import AutoTokenizer
        return cls._prepare_text_tokenizer(AutoTokenizer.from_pretrained(pretrained_model_path, **kwargs))

    @classmethod
    def init_image_tokenizer(cls, pretrained_model_path, offset, vision_config_overwrite=None):      # :153-165
        import os
        import yaml
        with open(os.path.join(pretrained_model_path, "vision_tokenizer_config.yaml")) as f:
            config = yaml.safe_load(f)
        config.update(vision_config_overwrite or {})
        params = config["params"]
        for key in ("ckpt_path",):
            if params.get(key):
                params[key] = os.path.join(pretrained_model_path, params[key])
        dd = params["ddconfig"]
        if dd.get("encoder_name"):
            dd["encoder_name"] = os.path.join(pretrained_model_path, dd["encoder_name"])
        return VisionTokenizer.from_config(config, token_offset=offset)

    def batch_decode(self, *a, **kw):
        return self.text_tokenizer.batch_decode(*a, **kw)

    @torch.no_grad()
    def forward(self, samples, **kwargs):
        """samples: {"language": [...], "vision": [...], ...} or a list of such dicts (tokenization_libra.py:167-175)."""
        from transformers import BatchEncoding
        if not isinstance(samples, (list, tuple)):
            samples = [samples]
        texts, images, ignore = [], [], []
        for sample in samples:
            for key, dst in (("language", texts), ("vision", images), ("contiguous_ignore_sign", ignore)):
                v = sample.get(key, None)
                if v is not None:
                    dst.extend(v) if isinstance(v, (list, tuple)) else dst.append(v)
        dev = self.device
        if images:
            images = [img.to(dev) for img in images]
            if images[0].dim() == 3:
                images = torch.stack(images)
            elif images[0].dim() == 4:
                images = torch.cat(images)
            else:
                raise ValueError("Invalid vision inputs.")
        else:
            images = None
        if ignore:
            ignore = torch.cat(ignore) if isinstance(ignore[0], torch.Tensor) else torch.tensor(ignore, device=dev)
        else:
            ignore = None
        has_image_flag = samples[-1].get("has_image", None)                   # the reference reads the LAST sample's flag (:211)
        if has_image_flag is not None:
            has_image_flag = torch.tensor(has_image_flag, device=dev, dtype=torch.bool)
        if not texts and images is None:
            raise ValueError("Empty inputs")
        if not texts:
            raise NotImplementedError
        if kwargs.pop("return_tensors", "pt") != "pt":
            raise ValueError("return_tensors = \"pt\" is fixed, and should not be specified to other values.")
        truncation = kwargs.pop("truncation", False)
        max_length = kwargs.pop("max_length", self.text_tokenizer.model_max_length)
        text_inputs = self.text_tokenizer(texts, return_tensors="pt", return_length=True, **kwargs)
        tt, it = self.text_tokenizer, self.image_tokenizer
        # the layout (which positions are image tokens, which are padding) is known on the host right here: hand it to the
        # model with the device tensors so that it never has to be copied back (schedule.attach_host_layout)
        ids_cpu, am_cpu = text_inputs["input_ids"], text_inputs["attention_mask"]
        flag_cpu = (ids_cpu == tt.img_ph_token_id) if images is not None else (ids_cpu == tt.img_gen_token_id)
        text_inputs = text_inputs.to(dev)
        input_ids = text_inputs["input_ids"]
        input_ids = torch.where(input_ids == tt.img_gen_token_id, torch.full_like(input_ids, it.boi_token_id), input_ids)
        gen_mask = text_inputs["input_ids"] == tt.img_gen_token_id
        image_ids = feat = None
        if images is not None:
            enc = it(images.to(self.dtype))
            image_ids, feat = enc["input_ids"], enc["encoder_feat"]
            if has_image_flag is not None:
                image_ids, feat = image_ids[:, has_image_flag], feat[has_image_flag]
        out = assemble_inputs(input_ids, text_inputs["attention_mask"], tt.img_ph_token_id, image_ids, feat,
                              max_vision_token_length=it.max_vision_token_length,
                              contiguous_ignore=None if ignore is None else ignore.to(dev), truncation=truncation,
                              max_length=max_length)
        if images is None:                                                        # generation prompts: <img_gen> opens an image (:274-275)
            out["vision_indices"] = torch.where(gen_mask[:, :out["vision_indices"].shape[1]],
                                                torch.zeros_like(out["vision_indices"]), out["vision_indices"])
        from ..schedule import attach_host_layout
        Tn = out["vision_indices"].shape[1]
        attach_host_layout(out["vision_indices"], flag_cpu[:, :Tn].contiguous())
        attach_host_layout(out["attention_mask"], am_cpu[:, :Tn].contiguous())
        return out if self.raw_output else BatchEncoding(out)
