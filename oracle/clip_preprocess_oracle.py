"""TEST INFRASTRUCTURE (oracle): CPU restatement, in numpy integer arithmetic, of the image preprocessing in front of the
vision tokenizer -- SURVEY.md section 8(f) N4, second half.  Only tests/, __graft_entry__.smoke() and bench.py's CPU leg
may import this file; the product path is csrc/preprocess.cu.

What it restates
  * libra/data/processors/libra_processor.py:44-60 (`Expand2Square`: paste the image centred on a square canvas of the
    processor's mean colour) and :65-84 / :96-111 (`LibraEvalImageProcessor` / `LibraImageProcessor`: the CLIP processor);
  * libra/models/clip/image_processing_clip.py:124-150 (resize: shortest edge -> size, other edge int(size * long / short)),
    :152-174 (center crop), :176-217 (rescale by 1/255, normalise by mean / std), :296-337 (the order of the steps);
  * the third-party code those call, absent from /root/reference: `transformers.image_transforms.{resize, center_crop,
    rescale, normalize}` (transformers==4.38.2 pinned in requirements.txt) and, below `resize`, Pillow's `Image.resize(...,
    resample=BICUBIC)`, i.e. `ImagingResample` (libImaging/Resample.c): separable two-pass convolution, horizontal pass
    first, coefficients computed in double precision (support 2.0 * max(scale, 1): antialiased when shrinking), normalised
    to sum 1, converted to 22-bit fixed point (rounded half away from zero), accumulated in int32 from 1 << 21, shifted
    right by 22 and clamped to uint8 after EACH pass.

Pinning (tests/test_preprocess_cpu.py): bit-exact against Pillow itself on the uint8 stage and exactly equal to the
reference's own CLIPImageProcessor (imported live through oracle/refshim.py where /root/reference exists, and through the
committed fixture tests/golden/clip_preprocess.pt elsewhere) on the float32 result.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def bicubic_filter(x: float) -> float:
    """Resample.c bicubic_filter, a = -0.5."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size).
    Returns bounds [out,2] (first input index, tap count), kk [out, ksize] int32, ksize."""
    support0 = 2.0
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size       # (double)(in1 - in0) / outSize with float in0, in1
    filterscale = scale if scale >= 1.0 else 1.0
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _resample_last_axis(img: np.ndarray, out_size: int) -> np.ndarray:
    """One 8-bit pass along axis 1 of img [rows, in_size, C] -> [rows, out_size, C] (ImagingResampleHorizontal_8bpc)."""
    rows, in_size, ch = img.shape
    bounds, kk, _ = precompute_coeffs(in_size, out_size)
    out = np.empty((rows, out_size, ch), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, x0:x0 + n, :], kk[xx, :n].astype(np.int64), axes=([1], [0]))
        # int32 wrap-around never happens for 8-bit data: |sum k| * 255 < 2^31; the shift is arithmetic (floor)
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bicubic(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Image.resize((out_w, out_h), resample=BICUBIC) of a uint8 [H, W, C] image: horizontal pass, then vertical pass
    (ImagingResample; a pass whose size does not change is skipped there and is the identity here)."""
    h, w, _ = img.shape
    tmp = _resample_last_axis(img, out_w) if out_w != w else img
    if out_h != h:
        tmp = _resample_last_axis(tmp.transpose(1, 0, 2), out_h).transpose(1, 0, 2)
    return np.ascontiguousarray(tmp)


def resize_output_size(h: int, w: int, shortest_edge: int) -> Tuple[int, int]:
    """transformers.image_transforms.get_resize_output_image_size(size=int, default_to_square=False) -> (height, width)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def expand2square(img: np.ndarray, background: Sequence[int]) -> np.ndarray:
    """libra_processor.py:44-60."""
    h, w, c = img.shape
    if w == h:
        return img
    s = max(w, h)
    out = np.empty((s, s, c), dtype=np.uint8)
    out[:] = np.asarray(background, dtype=np.uint8)
    if w > h:
        top = (w - h) // 2
        out[top:top + h] = img
    else:
        left = (h - w) // 2
        out[:, left:left + w] = img
    return out


def normalize_lut(mean: Sequence[float], std: Sequence[float], rescale_factor: float = 1 / 255) -> np.ndarray:
    """[3, 256] float32: rescale (float64 product, cast to float32) then (x - mean) / std in float32 -- the reference order
    (image_processing_clip.py:328-332 over transformers.image_transforms.rescale / normalize)."""
    v = (np.arange(256, dtype=np.uint8).astype(np.float64) * rescale_factor).astype(np.float32)
    m, s = np.array(mean, dtype=np.float32), np.array(std, dtype=np.float32)
    return ((v[None, :] - m[:, None]) / s[:, None]).astype(np.float32)


def clip_preprocess_u8(img: np.ndarray, size: int = 336, crop: int = 336, pad_square: Optional[Sequence[int]] = None) -> np.ndarray:
    """uint8 [H, W, 3] -> uint8 [crop, crop, 3]: (optional Expand2Square) -> resize shortest edge -> center crop."""
    if pad_square is not None:
        img = expand2square(img, pad_square)
    h, w, _ = img.shape
    oh, ow = resize_output_size(h, w, size)
    r = pil_resize_bicubic(img, ow, oh)
    top, left = (oh - crop) // 2, (ow - crop) // 2
    if top < 0 or left < 0:
        raise NotImplementedError("crop larger than the resized image (zero padding branch of center_crop)")
    return r[top:top + crop, left:left + crop]


def clip_preprocess(img: np.ndarray, size: int = 336, crop: int = 336, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD,
                    pad_square: Optional[Sequence[int]] = None) -> np.ndarray:
    """uint8 [H, W, 3] -> float32 [3, crop, crop] = CLIPImageProcessor(...)(img)["pixel_values"][0]."""
    u8 = clip_preprocess_u8(img, size, crop, pad_square)
    lut = normalize_lut(mean, std)
    return np.stack([lut[c][u8[:, :, c]] for c in range(3)])
