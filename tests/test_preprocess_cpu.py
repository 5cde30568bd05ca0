"""N4 (second half), CPU side: the numpy oracle of the CLIP image preprocessing (oracle/clip_preprocess_oracle.py) against
Pillow, against transformers' PIL-based CLIP processor (= the class the reference vendors, libra/models/clip/
image_processing_clip.py), against the live reference where /root/reference exists and against the committed fixture; and
the library's HOST coefficient tables against the oracle's (bit exact)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import clip_preprocess_oracle as O

SIZES = [(500, 733), (733, 500), (336, 336), (100, 80), (1200, 1600), (337, 900), (224, 224), (2000, 350), (33, 47), (336, 1000),
         (1, 1), (2, 3), (1, 50), (60, 2), (335, 337)]        # degenerate: one pixel, extreme aspect ratios, off-by-one sizes


def fixture_pixel_values(g, key, i):
    """tests/golden/clip_preprocess.pt stores the reference's float32 output as (uint8 index, the reference's own 3x256 table);
    oracle/make_golden.py asserted that the pair reproduces it bit for bit."""
    lut, idx = g["lut"].numpy(), g[key][i].numpy()
    return np.stack([lut[c][idx[..., c]] for c in range(3)])


def _img(h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if seed % 2:                                     # smooth content next to noise: exercises negative lobes + clamping differently
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.stack([(yy * 255 // max(h - 1, 1)), (xx * 255 // max(w - 1, 1)), ((yy + xx) % 256)], -1).astype(np.uint8)
    return base


@pytest.mark.parametrize("hw", SIZES)
def test_oracle_resize_is_bit_exact_with_pillow(hw):
    from PIL import Image
    h, w = hw
    img = _img(h, w, h + w)
    oh, ow = O.resize_output_size(h, w, 336)
    mine = O.pil_resize_bicubic(img, ow, oh)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
    assert np.array_equal(mine, ref)


@pytest.mark.parametrize("hw", SIZES[:6])
def test_oracle_equals_transformers_pil_processor(hw):
    tfm = pytest.importorskip("transformers")
    if not hasattr(tfm, "CLIPImageProcessorPil"):
        pytest.skip("no PIL-backed CLIP processor in this transformers")
    P = tfm.CLIPImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    img = _img(*hw, seed=3)
    want = P(img, return_tensors="np")["pixel_values"][0]
    got = O.clip_preprocess(img)
    assert got.dtype == np.float32 and np.array_equal(got, want)


def test_oracle_equals_live_reference_processor():
    from oracle import refshim
    if not refshim.reference_available():
        pytest.skip("/root/reference not present")
    refshim.install()
    import importlib
    from PIL import Image
    m = importlib.import_module("libra.models.clip.image_processing_clip")
    Expand2Square = refshim.load_reference_class("libra/data/processors/libra_processor.py", "Expand2Square",
                                                  {"torch": torch, "Image": Image})
    P = m.CLIPImageProcessor(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    bg = tuple(int(x * 255) for x in P.image_mean)
    for hw in [(500, 733), (90, 61), (700, 352)]:
        img = _img(*hw, seed=8)
        assert np.array_equal(O.clip_preprocess(img), P(img, return_tensors="np")["pixel_values"][0])
        sq = Expand2Square(bg)(Image.fromarray(img))                         # the eval processor's first step
        want = P(sq, return_tensors="np")["pixel_values"][0]
        assert np.array_equal(O.clip_preprocess(img, pad_square=bg), want)


def test_oracle_matches_committed_fixture(golden):
    g = golden("clip_preprocess")
    for i, img in enumerate(g["images"]):
        a = img.numpy()
        assert np.array_equal(O.clip_preprocess(a), fixture_pixel_values(g, "index", i)), i
        assert np.array_equal(O.clip_preprocess(a, pad_square=tuple(g["background"])), fixture_pixel_values(g, "index_square", i)), i


@pytest.mark.parametrize("in_size,out_size", [(733, 492), (500, 336), (80, 336), (1600, 448), (336, 336), (4000, 336), (47, 478)])
def test_library_host_tables_equal_the_oracle(in_size, out_size):
    from libra_b200 import _lib
    lib = _lib.load()
    bounds, kk, ksize = O.precompute_coeffs(in_size, out_size)
    first, n = (out_size - 336) // 2 if out_size >= 336 else 0, min(336, out_size)
    cb = (ctypes.c_int32 * (n * ksize))()
    bb = (ctypes.c_int32 * (2 * n))()
    got = lib.lb_clip_resample_coeffs(in_size, out_size, first, n, cb, n * ksize, bb)
    assert got == ksize
    assert np.array_equal(np.frombuffer(cb, dtype=np.int32).reshape(n, ksize), kk[first:first + n])
    assert np.array_equal(np.frombuffer(bb, dtype=np.int32).reshape(n, 2), bounds[first:first + n])


def test_normalize_lut_rounding_sequence():
    lut = O.normalize_lut(O.OPENAI_CLIP_MEAN, O.OPENAI_CLIP_STD)
    v = np.arange(256, dtype=np.uint8)
    x = (v.astype(np.float64) * (1 / 255)).astype(np.float32)
    for c in range(3):
        want = (x - np.float32(O.OPENAI_CLIP_MEAN[c])) / np.float32(O.OPENAI_CLIP_STD[c])
        assert np.array_equal(lut[c], want)


# ---------------------------------------------------------------------------------------------------------------------
# The indexing of the word-load kernels (csrc/preprocess.cu: resize_h_words_kernel / resize_v_words_kernel), restated thread by
# thread in Python: 4 taps = 12 source bytes fetched as 4 aligned 32-bit words and funnel-shifted to the window's first byte
# (16 bytes of slack behind the packed buffer), canvas colour x sum of the outside taps, and in the vertical pass one thread
# per 4 consecutive bytes of the output row.  Same integer sums as the oracle => the kernels' arithmetic is pinned on the CPU
# (the GPU tests then only have to show that the CUDA code is this algorithm).
PB = O.PRECISION_BITS


def _clip8(v):
    v >>= PB
    return 0 if v < 0 else (255 if v > 255 else v)


def _funnelshift_r(lo, hi, sh):
    return (((hi << 32) | lo) >> sh) & 0xFFFFFFFF


def _word(buf, byte_off):
    return int(buf[byte_off]) | int(buf[byte_off + 1]) << 8 | int(buf[byte_off + 2]) << 16 | int(buf[byte_off + 3]) << 24


def emulate(img, size=336, crop=336, pad=None):
    h, w, _ = img.shape
    ch, cw, pad_top, pad_left = h, w, 0, 0
    if pad is not None and h != w:                       # Expand2Square: a virtual canvas, never materialised
        if w > h:
            pad_top = (w - h) // 2
        else:
            pad_left = (h - w) // 2
        ch = cw = max(h, w)
    oh, ow = O.resize_output_size(ch, cw, size)
    top, left = (oh - crop) // 2, (ow - crop) // 2
    bh, kh, _ = O.precompute_coeffs(cw, ow)
    bv, kv, _ = O.precompute_coeffs(ch, oh)
    bh, kh = bh[left:left + crop], kh[left:left + crop]  # tables for the crop window only
    bv, kv = bv[top:top + crop], kv[top:top + crop]
    y_first = int(bv[0, 0])
    rows = int(bv[-1, 0] + bv[-1, 1]) - y_first
    flat = np.concatenate([img.reshape(-1), np.zeros(16, np.uint8)])       # the 16 bytes of slack the C ABI asks for
    bg = pad if pad is not None else (0, 0, 0)
    pitch = crop * 3
    tmp = np.zeros((rows, pitch), np.uint8)
    for xx in range(crop):                               # pass 1: thread = crop column
        x0, n, k = int(bh[xx, 0]) - pad_left, int(bh[xx, 1]), kh[xx]
        t0 = min(n, max(0, -x0))
        t1 = max(t0, min(n, w - x0))
        kb = int(k[:t0].sum() + k[t1:n].sum())           # taps on the canvas colour
        kall = kb + int(k[t0:t1].sum())
        for r in range(rows):
            cy = y_first + r - pad_top
            s = [1 << (PB - 1)] * 3
            if 0 <= cy < h:
                for c0 in range(0, t1 - t0, 4):          # 4 taps = 12 bytes = 4 aligned words
                    addr = (cy * w + (x0 + t0 + c0)) * 3
                    aligned, sh = addr & ~3, (addr & 3) * 8
                    wd = [_word(flat, aligned + 4 * i) for i in range(4)]
                    a = [_funnelshift_r(wd[i], wd[i + 1], sh) for i in range(3)]
                    for q in range(4):
                        t = t0 + c0 + q
                        kq = int(k[t]) if t < t1 else 0
                        for c in range(3):
                            bi = 3 * q + c
                            s[c] += ((a[bi >> 2] >> (8 * (bi & 3))) & 0xFF) * kq
                s = [s[c] + bg[c] * kb for c in range(3)]
            else:
                s = [s[c] + bg[c] * kall for c in range(3)]
            tmp[r, xx * 3:xx * 3 + 3] = [_clip8(v) for v in s]
    out = np.zeros((crop, pitch), np.uint8)
    for yy in range(crop):                               # pass 2: thread = 4 consecutive bytes of the output row
        y0, n = int(bv[yy, 0]) - y_first, int(bv[yy, 1])
        for j in range(pitch // 4):
            s = [1 << (PB - 1)] * 4
            for t in range(n):
                wv = _word(tmp[y0 + t], 4 * j)
                for b in range(4):
                    s[b] += ((wv >> (8 * b)) & 0xFF) * int(kv[yy, t])
            out[yy, 4 * j:4 * j + 4] = [_clip8(v) for v in s]
    return out.reshape(crop, crop, 3)


@pytest.mark.parametrize("case", [(40, 61, 24, None), (61, 40, 24, (122, 116, 104)), (24, 24, 24, None), (7, 5, 24, (1, 2, 3)),
                                  (100, 30, 24, None), (30, 100, 24, (9, 8, 7)), (50, 50, 28, None)])
def test_word_load_kernel_indexing_equals_the_oracle(case):
    h, w, size, pad = case
    img = np.random.default_rng(h * 131 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(emulate(img, size, size, pad), O.clip_preprocess_u8(img, size=size, crop=size, pad_square=pad))
