"""Image processors of the reference's data path on the GPU (SURVEY.md section 8f, N4 second half).

Mirrors
  * `CLIPImageProcessor` as the reference uses it (libra/models/clip/image_processing_clip.py:91-122 constructor fields,
    :219-337 `preprocess`; HF's class verbatim): convert to RGB -> resize shortest edge (Pillow BICUBIC) -> center crop ->
    rescale 1/255 -> normalise -> channels first;
  * `LibraImageProcessor` ("libra_image") and `LibraEvalImageProcessor` ("libra_image_eval", Expand2Square with the mean
    colour first) of libra/data/processors/libra_processor.py:65-111, registered under the same names.

The arithmetic runs in `lb_clip_preprocess` (csrc/preprocess.cu), bit-exact with Pillow's 8-bit resampler and with the
reference's float32 rounding sequence (tests/test_gpu_preprocess.py); decoding a compressed file into a uint8 array stays
with the caller (PIL on the host: the CPU data pipeline is out of scope).  A batch is packed into one pinned host buffer and
copied once; images that are already CUDA tensors are packed on the device.  There is no CPU fallback: without the library
and an sm_100 device the call raises.
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .ops import DT_BF16, DT_F32, _p, _st
from .registry import registry

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_BICUBIC = 3          # PIL.Image.Resampling.BICUBIC


def _edge(size, key_a, key_b=None) -> int:
    if isinstance(size, dict):
        if key_a in size:
            return int(size[key_a])
        if key_b is not None and key_b in size:
            return int(size[key_b])
        raise ValueError(f"size dict must contain '{key_a}', got {sorted(size)}")
    if isinstance(size, (tuple, list)):
        return int(size[0])
    return int(size)


def _to_u8_hwc(img) -> Union[np.ndarray, torch.Tensor]:
    """PIL image / numpy / torch -> uint8 [H, W, 3] (numpy for host inputs, the tensor itself for CUDA inputs)."""
    if isinstance(img, torch.Tensor):
        t = img
        if t.dtype != torch.uint8:
            raise TypeError("tensor images must be uint8 (the rescale step is part of the processor)")
        if t.dim() == 3 and t.shape[0] == 3 and t.shape[2] != 3:
            t = t.permute(1, 2, 0)
        if t.dim() != 3 or t.shape[2] != 3:
            raise ValueError(f"expected an RGB image [H, W, 3] or [3, H, W], got {tuple(img.shape)}")
        return t.contiguous() if t.is_cuda else np.ascontiguousarray(t.numpy())
    if hasattr(img, "convert") and hasattr(img, "size"):                 # PIL.Image: convert_to_rgb (:317-318)
        if img.mode != "RGB":
            img = img.convert("RGB")
        return np.ascontiguousarray(np.asarray(img, dtype=np.uint8))
    a = np.asarray(img)
    if a.dtype != np.uint8:
        raise TypeError("array images must be uint8 (the rescale step is part of the processor)")
    if a.ndim == 3 and a.shape[0] == 3 and a.shape[2] != 3:
        a = a.transpose(1, 2, 0)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError(f"expected an RGB image [H, W, 3] or [3, H, W], got {a.shape}")
    return np.ascontiguousarray(a)


class CLIPImageProcessor:
    """Drop-in for the reference's CLIPImageProcessor on its default path (do_resize, do_center_crop, do_rescale,
    do_normalize, do_convert_rgb all on, BICUBIC).  `preprocess(images, return_tensors="pt")` returns
    {"pixel_values": [n, 3, crop, crop]} on the CUDA device (`dtype` float32 like the reference, or bfloat16 to feed the
    vision tower directly)."""

    model_input_names = ["pixel_values"]

    def __init__(self, do_resize=True, size=None, resample=_BICUBIC, do_center_crop=True, crop_size=None, do_rescale=True,
                 rescale_factor=1 / 255, do_normalize=True, image_mean=None, image_std=None, do_convert_rgb=True,
                 pad_to_square: bool = False, background_color: Optional[Sequence[int]] = None, device=None,
                 dtype=torch.float32, **unused):
        if not (do_resize and do_center_crop and do_rescale and do_normalize and do_convert_rgb):
            raise NotImplementedError("the CUDA processor implements the reference's default pipeline: every step switched on")
        if int(resample) != _BICUBIC:
            raise NotImplementedError("resample must be PIL BICUBIC (3), the reference's setting")
        size = size if size is not None else {"shortest_edge": 224}
        crop_size = crop_size if crop_size is not None else {"height": 224, "width": 224}
        self.size = {"shortest_edge": _edge(size, "shortest_edge")}
        ch, cw = _edge(crop_size, "height"), _edge(crop_size, "width", "height")
        if ch != cw:
            raise NotImplementedError("square crops only")
        self.crop_size = {"height": ch, "width": cw}
        self.rescale_factor = float(rescale_factor)
        self.image_mean = tuple(float(x) for x in (image_mean if image_mean is not None else OPENAI_CLIP_MEAN))
        self.image_std = tuple(float(x) for x in (image_std if image_std is not None else OPENAI_CLIP_STD))
        self.pad_to_square = bool(pad_to_square)
        bg = background_color if background_color is not None else tuple(int(x * 255) for x in self.image_mean)
        self.background_color = tuple(int(x) for x in bg)              # libra_processor.py:74-75
        self.device = device
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("dtype must be float32 or bfloat16")
        self.dtype = dtype

    @classmethod
    def from_pretrained(cls, path, **kwargs):
        """Reads HF's preprocessor_config.json from a checkpoint directory (what CLIPImageProcessor.from_pretrained does)."""
        f = path if os.path.isfile(path) else os.path.join(path, "preprocessor_config.json")
        with open(f, "r") as fh:
            cfg = json.load(fh)
        for k in ("image_processor_type", "feature_extractor_type", "processor_class"):
            cfg.pop(k, None)
        cfg.update(kwargs)
        return cls(**cfg)

    def to_dict(self) -> Dict:
        return {"do_resize": True, "size": dict(self.size), "resample": _BICUBIC, "do_center_crop": True,
                "crop_size": dict(self.crop_size), "do_rescale": True, "rescale_factor": self.rescale_factor, "do_normalize": True,
                "image_mean": list(self.image_mean), "image_std": list(self.image_std), "do_convert_rgb": True,
                "image_processor_type": "CLIPImageProcessor"}

    # ------------------------------------------------------------------------------------------------------------------
    def pack(self, images):
        """uint8 images (PIL / numpy / torch, HWC or CHW) -> (packed device buffer, byte offsets, heights, widths): one pinned
        staging buffer and ONE host->device copy for host inputs, device-side packing for CUDA tensors."""
        _lib.require_device()
        single = not isinstance(images, (list, tuple)) and not (isinstance(images, (np.ndarray, torch.Tensor)) and images.ndim == 4)
        items = [images] if single else list(images)
        arrs = [_to_u8_hwc(im) for im in items]
        n = len(arrs)
        if n == 0:
            raise ValueError("no images")
        dev = torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        hs = (ctypes.c_int32 * n)(*[int(a.shape[0]) for a in arrs])
        ws = (ctypes.c_int32 * n)(*[int(a.shape[1]) for a in arrs])
        sizes = [int(a.shape[0]) * int(a.shape[1]) * 3 for a in arrs]
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot)
            tot += (s + 15) // 16 * 16
        offsets = (ctypes.c_int64 * n)(*offs)
        tot += 16                                           # lb_clip_preprocess may read 16 bytes past the last image
        with torch.cuda.device(dev):
            if all(isinstance(a, torch.Tensor) for a in arrs):
                packed = torch.empty(tot, dtype=torch.uint8, device=dev)
                for a, o, s in zip(arrs, offs, sizes):
                    packed[o:o + s].copy_(a.to(dev).reshape(-1))
            else:
                host = torch.empty(tot, dtype=torch.uint8).pin_memory()
                hv = host.numpy()
                for a, o, s in zip(arrs, offs, sizes):
                    hv[o:o + s] = (a.cpu().numpy() if isinstance(a, torch.Tensor) else a).reshape(-1)
                packed = host.to(dev, non_blocking=True)
        return packed, offsets, hs, ws

    def run_packed(self, packed, offsets, hs, ws, return_uint8: bool = False):
        """The C-ABI call on an already packed batch: [n, 3, crop, crop] pixel_values (+ the uint8 stage on request)."""
        n = len(hs)
        dev = packed.device
        with torch.cuda.device(dev):
            size, crop = self.size["shortest_edge"], self.crop_size["height"]
            lib = _lib.load()
            need = int(lib.lb_clip_preprocess_workspace(hs, ws, n, size, crop, int(self.pad_to_square)))
            if need < 0:
                raise _lib.LibraB200Error(f"lb_clip_preprocess_workspace failed: {_lib.last_error()}")
            work = torch.empty(need, dtype=torch.uint8, device=dev)
            out = torch.empty(n, 3, crop, crop, dtype=self.dtype, device=dev)
            u8 = torch.empty(n, crop, crop, 3, dtype=torch.uint8, device=dev) if return_uint8 else None
            bg = (ctypes.c_uint8 * 3)(*self.background_color)
            mean = (ctypes.c_float * 3)(*self.image_mean)
            std = (ctypes.c_float * 3)(*self.image_std)
            _lib.call("lb_clip_preprocess", _p(packed), offsets, hs, ws, n, size, crop, int(self.pad_to_square), bg, mean, std,
                      ctypes.c_double(self.rescale_factor), _p(out), DT_BF16 if self.dtype == torch.bfloat16 else DT_F32,
                      _p(u8), _p(work), need, _st())
            # `packed` / `work` go back to the caching allocator in stream order: safe for later work on this stream
        return out, u8

    def preprocess(self, images, return_tensors: Optional[str] = "pt", return_uint8: bool = False, **unused):
        if return_tensors not in (None, "pt"):
            raise NotImplementedError("the CUDA processor returns torch tensors (return_tensors='pt')")
        out, u8 = self.run_packed(*self.pack(images), return_uint8=return_uint8)
        data = {"pixel_values": out}
        if return_uint8:
            data["uint8"] = u8
        return data

    __call__ = preprocess


@registry.register_processor("libra_image")
class LibraImageProcessor:
    """libra_processor.py:96-117: the CLIP processor of a checkpoint directory; `__call__(item)` returns pixel_values[0]."""

    def __init__(self, pretrained_path=None, processor: Optional[CLIPImageProcessor] = None, **kwargs):
        self.transform = processor if processor is not None else self.build_transforms(pretrained_path, **kwargs)

    @classmethod
    def build_transforms(cls, pretrained_path, **kwargs):
        return CLIPImageProcessor.from_pretrained(pretrained_path, **kwargs)

    def __call__(self, item, image_size=None):
        return self.transform(item, return_tensors="pt")["pixel_values"][0]

    @classmethod
    def from_config(cls, cfg=None):
        path = cfg.get("pretrained_path", None) if cfg is not None else None
        return cls(pretrained_path=path)


@registry.register_processor("libra_image_eval")
class LibraEvalImageProcessor(LibraImageProcessor):
    """libra_processor.py:65-93: Expand2Square(mean colour) in front of the CLIP processor."""

    @classmethod
    def build_transforms(cls, pretrained_path, **kwargs):
        return CLIPImageProcessor.from_pretrained(pretrained_path, pad_to_square=True, **kwargs)

    def __init__(self, pretrained_path=None, processor: Optional[CLIPImageProcessor] = None, **kwargs):
        if processor is not None and not processor.pad_to_square:
            raise ValueError("the eval processor pads to a square first (Expand2Square)")
        super().__init__(pretrained_path, processor, **kwargs)
