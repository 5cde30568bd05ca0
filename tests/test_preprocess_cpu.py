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
