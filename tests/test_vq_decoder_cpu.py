"""N2 host side (no GPU): module tree / state-dict keys of the decode side against the fixture the reference produced, the
nearest-upsample index rule against F.interpolate, weight packing for the nine-segment convolution."""
import torch
import torch.nn.functional as F


def test_state_dict_keys_match_reference_fixture(golden):
    from libra_b200.models.vq_decoder import VQDecoder
    g = golden("vq_decode_tiny")
    sd = g["state_dict"]
    m = VQDecoder(dict(g["config"]), embed_dim=sd["post_quant_conv.weight"].shape[1], codebook_size=g["codebook_size"],
                  num_codebook=g["ids"].shape[0], token_offset=g["token_offset"])
    own = m.state_dict()
    assert set(own) == set(sd)
    assert all(tuple(own[k].shape) == tuple(sd[k].shape) for k in sd)
    m.load_state_dict(sd, strict=True)
    assert not any(p.requires_grad for p in m.parameters()) and not m.training
    assert m.boi_token_id == g["boi_token_id"]


def test_nearest_source_index_is_interpolates_rule():
    from libra_b200.models.vq_decoder import nearest_source_index
    for n, s in ((6, 2.0), (12, 4.0), (24, 336 / 96), (7, 1.5), (96, 3.5), (5, 1.0)):
        x = torch.arange(n, dtype=torch.float32).view(1, 1, n, 1).expand(1, 1, n, 2)
        want = F.interpolate(x, scale_factor=s, mode="nearest")[0, 0, :, 0].to(torch.int32)
        got = nearest_source_index(n, s)
        assert torch.equal(got, want), (n, s)


def test_pack_layout_cpu(golden):
    """[Co,Ci,3,3] -> [9,Co8,Ci8]: tap t = 3*dy + dx holds weight[:, :, dy, dx]; out_ch 3 is zero-padded to 8."""
    from libra_b200.models.vq_decoder import VQDecoder
    g = golden("vq_decode_tiny")
    sd = g["state_dict"]
    m = VQDecoder(dict(g["config"]), embed_dim=sd["post_quant_conv.weight"].shape[1], codebook_size=g["codebook_size"],
                  num_codebook=g["ids"].shape[0], token_offset=g["token_offset"])
    m.load_state_dict(sd)
    P = m._pack()
    w = sd["decoder.conv_out.weight"]
    t = P["decoder.conv_out.w"]
    assert tuple(t.shape) == (9, 8, w.shape[1])
    assert torch.equal(t[5, :3].float(), w[:, :, 1, 2].bfloat16().float()) and bool((t[:, 3:] == 0).all())
    assert tuple(P["post_quant_conv.w"].shape) == (32, 24) and tuple(P["quantize.project_out.w"].shape) == (24, 24)
    assert bool((P["quantize.project_out.w"][:, 18:] == 0).all())
