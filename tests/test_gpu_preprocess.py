"""N4 (second half) on the GPU: lb_clip_preprocess (csrc/preprocess.cu) through libra_b200.processors against the numpy oracle,
against Pillow + transformers' PIL-backed CLIP processor (the class the reference vendors) and against the fixture the live
reference produced.  Integer stage bit exact, float32 result exactly equal (the arithmetic is a table of 256 values per channel)."""
import numpy as np
import pytest
import torch

from gpu_util import need_gpu
from oracle import clip_preprocess_oracle as O
from test_preprocess_cpu import _img, fixture_pixel_values

pytestmark = pytest.mark.gpu
SIZE = dict(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})


@pytest.mark.parametrize("pad", [False, True])
def test_batch_of_mixed_sizes_equals_oracle(pad):
    need_gpu()
    from libra_b200.processors import CLIPImageProcessor
    shapes = [(500, 733), (733, 500), (336, 336), (100, 80), (1200, 1600), (337, 900), (224, 224), (2000, 350), (33, 47), (336, 1000),
              (1, 1), (2, 3), (1, 50), (60, 2), (335, 337)]
    imgs = [_img(h, w, h * 3 + w) for h, w in shapes]
    P = CLIPImageProcessor(pad_to_square=pad, **SIZE)
    out = P(imgs, return_tensors="pt", return_uint8=True)
    pv, u8 = out["pixel_values"].cpu().numpy(), out["uint8"].cpu().numpy()
    assert pv.shape == (len(imgs), 3, 336, 336) and pv.dtype == np.float32
    bg = P.background_color if pad else None
    for i, im in enumerate(imgs):
        assert np.array_equal(u8[i], O.clip_preprocess_u8(im, pad_square=bg)), (i, shapes[i])
        assert np.array_equal(pv[i], O.clip_preprocess(im, pad_square=bg)), (i, shapes[i])


@pytest.mark.parametrize("size", [57, 224])
def test_other_target_sizes_and_the_byte_kernels(size):
    """crop % 4 != 0 takes the byte-per-load kernels, crop % 4 == 0 the word-load ones (csrc/preprocess.cu): both against the oracle."""
    need_gpu()
    from libra_b200.processors import CLIPImageProcessor
    imgs = [_img(h, w, h + 7 * w) for h, w in [(300, 200), (64, 64), (90, 333), (5, 9)]]
    for pad in (False, True):
        P = CLIPImageProcessor(size={"shortest_edge": size}, crop_size={"height": size, "width": size}, pad_to_square=pad)
        out = P(imgs, return_uint8=True)
        bg = P.background_color if pad else None
        for i, im in enumerate(imgs):
            assert np.array_equal(out["uint8"][i].cpu().numpy(), O.clip_preprocess_u8(im, size=size, crop=size, pad_square=bg)), (i, pad)
            assert np.array_equal(out["pixel_values"][i].cpu().numpy(), O.clip_preprocess(im, size=size, crop=size, pad_square=bg)), (i, pad)


def test_equals_pillow_and_transformers_pil_processor():
    need_gpu()
    tfm = pytest.importorskip("transformers")
    from PIL import Image
    from libra_b200.processors import CLIPImageProcessor
    imgs = [_img(480, 640, 1), _img(3000, 4000, 2), _img(640, 427, 3)]
    P = CLIPImageProcessor(**SIZE)
    out = P([Image.fromarray(a) for a in imgs], return_uint8=True)
    u8 = out["uint8"].cpu().numpy()
    for i, a in enumerate(imgs):
        h, w = a.shape[:2]
        oh, ow = O.resize_output_size(h, w, 336)
        ref = np.asarray(Image.fromarray(a).resize((ow, oh), resample=Image.BICUBIC))
        t, l = (oh - 336) // 2, (ow - 336) // 2
        assert np.array_equal(u8[i], ref[t:t + 336, l:l + 336]), i
    if hasattr(tfm, "CLIPImageProcessorPil"):
        H = tfm.CLIPImageProcessorPil(**SIZE)
        want = H(imgs, return_tensors="np")["pixel_values"]
        assert np.array_equal(out["pixel_values"].cpu().numpy(), np.stack(want))


def test_reference_fixture_and_eval_processor(golden):
    need_gpu()
    from libra_b200.processors import CLIPImageProcessor, LibraEvalImageProcessor, LibraImageProcessor
    from libra_b200.registry import registry
    g = golden("clip_preprocess")
    imgs = [t.numpy() for t in g["images"]]
    plain = LibraImageProcessor(processor=CLIPImageProcessor(**SIZE))
    ev = registry.get_processor_class("libra_image_eval")(processor=CLIPImageProcessor(pad_to_square=True, **SIZE))
    assert isinstance(ev, LibraEvalImageProcessor) and ev.transform.background_color == tuple(g["background"])
    for i, a in enumerate(imgs):
        assert np.array_equal(plain(a).cpu().numpy(), fixture_pixel_values(g, "index", i)), i
        assert np.array_equal(ev(a).cpu().numpy(), fixture_pixel_values(g, "index_square", i)), i


def test_input_forms_and_bf16_output():
    need_gpu()
    from PIL import Image
    from libra_b200.processors import CLIPImageProcessor
    a = _img(300, 410, 9)
    P = CLIPImageProcessor(**SIZE)
    want = P(a)["pixel_values"]
    assert want.is_cuda and want.shape == (1, 3, 336, 336)
    for form in (torch.from_numpy(a), torch.from_numpy(a).cuda(), torch.from_numpy(a).permute(2, 0, 1).contiguous().cuda(),
                 Image.fromarray(a), Image.fromarray(a).convert("RGBA"), [a]):
        assert torch.equal(P(form)["pixel_values"], want)
    Pb = CLIPImageProcessor(dtype=torch.bfloat16, **SIZE)
    assert torch.equal(Pb(a)["pixel_values"], want.to(torch.bfloat16))
    with pytest.raises(TypeError):
        P(a.astype(np.float32))
    with pytest.raises(NotImplementedError):
        CLIPImageProcessor(do_center_crop=False)


def test_pixels_feed_the_vision_tokenizer():
    """End of the widened path: uint8 images -> lb_clip_preprocess (bf16) -> VisionTokenizer.encode gives the same ids as the
    reference-order pipeline (float32 pixel_values from the oracle, cast to bf16)."""
    need_gpu()
    from libra_b200.models import CLIPVisionConfig
    from libra_b200.models.tokenization_libra import VisionTokenizer
    from libra_b200.processors import CLIPImageProcessor
    cfg = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=56, patch_size=14)
    torch.manual_seed(0)
    vt = VisionTokenizer(cfg, select_layer=(-2, -3), embed_dim=18, token_offset=1000).to(torch.bfloat16).cuda()
    imgs = [_img(90, 61, 1), _img(200, 300, 2)]
    P = CLIPImageProcessor(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56}, dtype=torch.bfloat16)
    ids = vt.encode(P(imgs)["pixel_values"])["input_ids"]
    ref_px = torch.from_numpy(np.stack([O.clip_preprocess(a, size=56, crop=56) for a in imgs])).to(torch.bfloat16).cuda()
    assert torch.equal(ids, vt.encode(ref_px)["input_ids"])
