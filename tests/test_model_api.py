"""API-surface parity of the host mirror (no GPU): class names, constructor, state-dict keys and shapes equal to the
reference's, checked against the live reference when present and against the golden fixture's key list always."""
import pytest
import torch

from oracle import refshim


def _tiny_cfg():
    from oracle.make_golden import TINY_LIBRA
    return dict(TINY_LIBRA)


def test_state_dict_matches_golden_keys(golden):
    from libra_b200.models import LibraConfig, LibraForCausalLM
    g = golden("decoder_tiny")
    model = LibraForCausalLM(LibraConfig(**g["config"]))
    sd = model.state_dict()
    for k, v in g["state_dict"].items():
        assert k in sd, k
        assert tuple(sd[k].shape) == tuple(v.shape), k
    extra = [k for k in sd if k not in g["state_dict"] and "placeholder" not in k and "inv_freq" not in k]
    assert not extra, extra
    missing, unexpected = model.load_state_dict(g["state_dict"], strict=False)
    assert not unexpected
    assert all(("placeholder" in k or "inv_freq" in k) for k in missing), missing


@pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")
def test_state_dict_matches_live_reference():
    from libra_b200.models import LibraConfig, LibraForCausalLM
    ref = refshim.import_reference()
    cfg = _tiny_cfg()
    theirs = ref.modeling_libra.LibraForCausalLM(ref.configuration_libra.LibraConfig(**cfg)).state_dict()
    ours = LibraForCausalLM(LibraConfig(**cfg)).state_dict()
    assert set(ours) == set(theirs)
    for k in theirs:
        assert tuple(ours[k].shape) == tuple(theirs[k].shape), k
        assert ours[k].dtype == theirs[k].dtype, k


def test_config_defaults_match_reference_values():
    from libra_b200.models import LibraConfig
    c = LibraConfig()
    assert (c.hidden_size, c.intermediate_size, c.num_hidden_layers, c.num_attention_heads, c.vocab_size) == (4096, 11008, 32, 32, 32000)
    assert (c.vision_down_ratio, c.vision_vocab_size, c.vision_codebook_num, c.max_vision_token_length) == (4, 514, 2, 578)
    assert (c.contiguous_signal_size, c.bridge_rank, c.use_bridge, c.concat_signals, c.norm_signals) == (2048, 8, True, True, True)
    assert c.model_type == "libra" and c.unsupported_branches() == []


def test_unsupported_branches_are_refused():
    from libra_b200.models import LibraConfig, LibraForCausalLM
    with pytest.raises(NotImplementedError):
        LibraForCausalLM(LibraConfig(**{**_tiny_cfg(), "use_2d_rope": True}))


def test_init_scheme_matches_reference_rules():
    from libra_b200.models import LibraConfig, LibraForCausalLM
    m = LibraForCausalLM(LibraConfig(**_tiny_cfg()))
    a = m.model.layers[0].self_attn
    assert float(a.vision_k_bridge_on_language.weight_B.abs().max()) == 0.0      # rank LibraLinear: B = 0 (:506-507)
    assert float(a.vision_q_proj.weight_B.abs().max()) > 0.0
    assert abs(float(a.q_proj.weight.std()) - 0.02) < 0.005


def test_forward_requires_cuda():
    from libra_b200 import _lib
    from libra_b200.models import LibraConfig, LibraForCausalLM
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = LibraForCausalLM(LibraConfig(**_tiny_cfg()))
    with pytest.raises(_lib.LibraB200Error):
        m(input_ids=torch.zeros(2, 1, 4, dtype=torch.long), vision_indices=torch.full((1, 4), 578))


def test_routing_and_work_lists():
    from libra_b200 import schedule
    flag = torch.zeros(2, 300, dtype=torch.bool)
    flag[0, 1:200] = True
    flag[1, 130:256] = True
    rt = schedule.build_routing(flag)
    assert rt.n_vis == 199 + 126 and rt.n_lang == 600 - rt.n_vis
    f = flag.reshape(-1)
    assert torch.equal(f[rt.perm.long()].to(torch.uint8), rt.flag_sorted)
    assert torch.equal(rt.perm[rt.inv.long()], torch.arange(600, dtype=torch.int32))
    assert not rt.flag_sorted[:rt.n_lang].any() and rt.flag_sorted[rt.n_lang:].all()
    w = schedule.build_attn_work(flag, 2, 300, True, "cpu", kv_end=[300, 280])
    items = {tuple(r[:3].tolist()) for r in w.work_q}
    # sample 0: tile0 mixed, tile1 mixed (vision to 199, language after), tile2 language
    assert {(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (0, 2, 0)} <= items and (0, 2, 1) not in items
    # sample 1: tile0 language only; tile1 (128..255) mixed: 128,129 language
    assert (1, 0, 1) not in items and {(1, 1, 0), (1, 1, 1)} <= items
    assert w.qtile_has.shape == (2, 2, 3)
    kv = {tuple(r.tolist()) for r in w.work_kv}
    assert (0, 2, 0, 2) in kv and (0, 0, 1, 0) in kv
    # heaviest first
    n_kv = [min(int(r[1]) + 1, 3) for r in w.work_q]
    assert n_kv == sorted(n_kv, reverse=True)
