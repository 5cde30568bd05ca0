"""Oracle (oracle/libra_oracle.py) vs the fixtures produced by RUNNING THE REFERENCE
(oracle/make_golden.py).  CPU only; pins the oracle wherever the suite runs."""
import torch

from oracle import libra_oracle as O


def test_decoder_tiny_forward_backward(golden):
    g = golden("decoder_tiny")
    d = O.LibraDims.from_config(g["config"])
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["state_dict"].items()}
    inp = g["inputs"]
    out = O.libra_forward(sd, d, inp["input_ids"], inp["vision_indices"], attention_mask=inp["attention_mask"],
                          contiguous_signal=inp["contiguous_signal"], labels=inp["labels"], return_hidden=True)
    assert torch.allclose(out["loss"], g["loss"], atol=1e-6, rtol=0)
    la = out["logits"][:, :, g["logits_positions"]]
    fin = torch.isfinite(g["logits_at"])
    assert torch.equal(fin, torch.isfinite(la))
    assert (la[fin] - g["logits_at"][fin]).abs().max() < 2e-6
    assert (torch.logsumexp(out["logits"], -1) - g["logits_lse"]).abs().max() < 5e-6
    assert (out["hidden_states"][1] - g["hidden_after_layer0"]).abs().max() < 2e-6
    out["loss"].backward()
    for n, gr in g["grads"].items():
        assert sd[n].grad is not None, n
        err = (sd[n].grad - gr).abs().max().item()
        assert err <= 1e-6 + 1e-4 * gr.abs().max().item(), (n, err)


def test_attention_hd128(golden):
    g = golden("attention_hd128")
    d = O.LibraDims.from_config(g["config"])
    sd = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    x = g["x"].clone().requires_grad_(True)
    B, T, _ = x.shape
    pos = torch.arange(T)[None].expand(B, T)
    y = O.attention_block(sd, "self_attn", d, x, g["flag"], pos, g["attention_mask"])
    assert (y - g["out"]).abs().max() < 2e-5
    (y * g["grad_out"]).sum().backward()
    assert (x.grad - g["grad_x"]).abs().max() < 1e-4
    for key, name in (("grad_kbridge_B", "self_attn.vision_k_bridge_on_language.weight_B"),
                      ("grad_vbridge_A", "self_attn.vision_v_bridge_on_vision.weight_A"),
                      ("grad_vq_A", "self_attn.vision_q_proj.weight_A")):
        ref = g[key]
        assert (sd[name].grad - ref).abs().max() <= 1e-5 + 1e-4 * ref.abs().max(), key


def test_clip_tiny(golden):
    g = golden("clip_tiny")
    c = O.ClipDims(**{k: v for k, v in g["config"].items() if k != "num_channels"})
    hs = O.clip_vision_hidden_states(g["state_dict"], c, g["pixel_values"])
    assert len(hs) == len(g["hidden_states"])
    for a, b in zip(hs, g["hidden_states"]):
        assert (a - b).abs().max() < 3e-5
    f = O.clip_tower_features(hs, [-2, -3])
    assert f.shape == (3, 16, 256)


def test_lfq_bit_exact(golden):
    g = golden("lfq")
    x = g["x"].permute(0, 2, 3, 1)                     # [B,24,24,18]
    idx = O.lfq_indices(x, 2, 9)
    assert torch.equal(idx, g["indices"].to(torch.int64))
    codes = O.lfq_codes(idx, 9).permute(0, 3, 1, 2)
    assert torch.equal(codes, g["codes"])
    assert torch.equal(codes, g["quantized"])
    xp = g["xp"].permute(0, 2, 3, 1)
    h = torch.nn.functional.linear(xp, g["proj_w"], g["proj_b"])
    assert torch.equal(O.lfq_indices(h, 2, 9), g["indices_p"].to(torch.int64))


def test_rmsnorm_rope(golden):
    g = golden("norms_rope")
    assert torch.equal(O.rmsnorm(g["x"], g["w"], 1e-6), g["y32"])
    assert torch.equal(O.rmsnorm(g["x"].bfloat16(), g["w"], 1e-6), g["ybf"])
    cos, sin = O.rope_tables(128, 2048, 10000.0, torch.float32, "cpu")
    pos = torch.arange(40)[None]
    c, s = cos[pos][:, None], sin[pos][:, None]
    assert torch.equal(O.rope_apply(g["q"], c, s), g["q_rot"])
    assert torch.equal(O.rope_apply(g["k"], c, s), g["k_rot"])
    assert torch.equal(O.rope_apply(g["k"] * 2, c, s), g["k2_rot"])


def test_oracle_vq_decode_matches_reference_golden(golden):
    """N2 oracle (ids -> pixels) against the fixture the reference's taming Decoder produced (oracle/make_golden.py)."""
    g = golden("vq_decode_tiny")
    cfg = g["config"]
    d = O.VQDecoderDims(**{k: v for k, v in cfg.items() if k not in ("dropout", "in_channels")})
    got = O.vq_decode(g["state_dict"], d, g["ids"], g["token_offset"], g["codebook_size"], boi_token_id=g["boi_token_id"])
    assert got.shape == g["pixels"].shape and (got - g["pixels"]).abs().max() < 2e-5
    code = g["ids"][:, :, 1:-1].reshape(2, 2, 6, 6).permute(1, 2, 3, 0) - g["token_offset"]
    z = O.lfq_indices_to_codes(code, 9)
    import torch.nn.functional as F
    z = F.linear(z.permute(0, 2, 3, 1), g["state_dict"]["quantize.project_out.weight"], g["state_dict"]["quantize.project_out.bias"]).permute(0, 3, 1, 2)
    assert (z - g["codes"]).abs().max() < 1e-6
