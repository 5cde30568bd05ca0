"""End-to-end LibraForCausalLM (CUDA path, bf16) vs the fixture produced by the REFERENCE in fp32
(tests/golden/decoder_tiny.pt) and vs the oracle on other layouts.

Tolerances: the model runs in bf16 like the reference's training dtype; BASELINE.md asks logits within 1e-3 rel of the
reference *measured against fp32 with the bf16 reference's own distance as the noise floor*.  We therefore assert
  err(ours_bf16, ref_fp32) <= 1.1 * err(oracle_bf16, ref_fp32) + 3e-4
on relative Frobenius errors (FACTOR / SLACK below, set from the measured values in profiles/r02_parity_tiny.jsonl), plus
absolute caps on the loss."""
import pytest
import torch

from gpu_util import need_gpu, rel_err
from oracle import libra_oracle as O

pytestmark = pytest.mark.gpu
dev = "cuda"



def _record(name, **vals):
    """LB_PARITY_LOG=<file>: append the measured errors of a parity assertion (profiles/r02_parity_tiny.jsonl was made this way)."""
    import json, os
    path = os.environ.get("LB_PARITY_LOG")
    if path:
        with open(path, "a") as fh:
            fh.write(json.dumps({"test": name, **{k: float(v) for k, v in vals.items()}}) + "\n")


# "No worse than the reference's own bf16 path": relative Frobenius error against fp32 within FACTOR x the bf16 oracle's +
# SLACK.  Measured on B200 (profiles/r02_parity_tiny.jsonl): ours 5.721e-3 / 5.401e-3 / 5.475e-3 against 5.738e-3 / 5.401e-3 /
# 5.476e-3 for the bf16 oracle (the two round at the same points), so the round-1 bound of 1.5x + 5e-3 is tightened to:
FACTOR, SLACK = 1.1, 3e-4

def build(g):
    from libra_b200.models import LibraConfig, LibraForCausalLM
    model = LibraForCausalLM(LibraConfig(**g["config"]))
    model.load_state_dict(g["state_dict"], strict=False)
    return model.to(torch.bfloat16).to(dev)


def oracle_bf16(g, inp, labels=True):
    sd = {k: (v.to(dev).bfloat16() if v.is_floating_point() else v.to(dev)) for k, v in g["state_dict"].items()}
    d = O.LibraDims.from_config(g["config"])
    return O.libra_forward(sd, d, inp["input_ids"].to(dev), inp["vision_indices"].to(dev),
                           attention_mask=inp["attention_mask"].to(dev),
                           contiguous_signal=inp["contiguous_signal"].to(dev).bfloat16(),
                           labels=inp["labels"].to(dev) if labels else None, return_hidden=True)


def test_forward_logits_and_loss_vs_reference_golden(golden):
    need_gpu()
    g = golden("decoder_tiny")
    inp = g["inputs"]
    model = build(g).eval()
    with torch.no_grad():
        out = model(input_ids=inp["input_ids"].to(dev), attention_mask=inp["attention_mask"].to(dev),
                    vision_indices=inp["vision_indices"].to(dev), contiguous_signal=inp["contiguous_signal"].to(dev),
                    labels=inp["labels"].to(dev), output_hidden_states=True)
        orc = oracle_bf16(g, inp)
    assert out.logits.shape == (2, 2, inp["input_ids"].shape[2], 320 + 514)
    sel = g["logits_positions"].to(dev)
    want = g["logits_at"].to(dev)
    got = out.logits[:, :, sel].float()
    fin = torch.isfinite(want)
    assert torch.equal(fin, torch.isfinite(got)), "the -inf placeholder pattern must match the reference"
    valid = inp["attention_mask"].to(dev)[:, sel].bool()[None, :, :, None].expand_as(fin) & fin
    e_ours = rel_err(got[valid], want[valid])
    e_orc = rel_err(orc["logits"][:, :, sel].float()[valid], want[valid])
    # hidden states: [B,T,C] in original order
    h1 = out.hidden_states[1].float()
    m = inp["attention_mask"].to(dev).bool()
    eh_ours = rel_err(h1[m], g["hidden_after_layer0"].to(dev)[m])
    eh_orc = rel_err(orc["hidden_states"][1].float()[m], g["hidden_after_layer0"].to(dev)[m])
    _record("golden_forward", logits_ours=e_ours, logits_oracle_bf16=e_orc, hidden_ours=eh_ours, hidden_oracle_bf16=eh_orc,
            loss_ours=abs(float(out.loss) - float(g["loss"])), loss_oracle_bf16=abs(float(orc["loss"]) - float(g["loss"])))
    assert e_ours <= FACTOR * e_orc + SLACK, (e_ours, e_orc)
    assert abs(float(out.loss) - float(g["loss"])) <= abs(float(orc["loss"]) - float(g["loss"])) + 2e-3      # measured 4.8e-6 vs 7.7e-3
    assert eh_ours <= FACTOR * eh_orc + SLACK, (eh_ours, eh_orc)
    lse = torch.logsumexp(out.logits.float(), -1)
    mm = m[None].expand_as(lse)
    assert rel_err(lse[mm], g["logits_lse"].to(dev)[mm]) < 2e-2


def test_training_step_gradients_vs_reference_golden(golden):
    need_gpu()
    g = golden("decoder_tiny")
    inp = g["inputs"]
    model = build(g).train()
    out = model(input_ids=inp["input_ids"].to(dev), attention_mask=inp["attention_mask"].to(dev),
                vision_indices=inp["vision_indices"].to(dev), contiguous_signal=inp["contiguous_signal"].to(dev),
                labels=inp["labels"].to(dev))
    assert out.logits is None            # fused loss: the 4.3 GB logits tensor is not materialised in training
    assert abs(float(out.loss) - float(g["loss"])) < 5e-2
    out.loss.backward()
    params = dict(model.named_parameters())
    # noise floor: the oracle in bf16 with autograd
    sd = {k: (v.to(dev).bfloat16().requires_grad_(True) if v.is_floating_point() else v.to(dev)) for k, v in g["state_dict"].items()}
    d = O.LibraDims.from_config(g["config"])
    o = O.libra_forward(sd, d, inp["input_ids"].to(dev), inp["vision_indices"].to(dev), attention_mask=inp["attention_mask"].to(dev),
                        contiguous_signal=inp["contiguous_signal"].to(dev).bfloat16(), labels=inp["labels"].to(dev))
    o["loss"].backward()
    worst = 0.0
    for n, want in g["grads"].items():
        want = want.to(dev)
        assert params[n].grad is not None, n
        e_ours = rel_err(params[n].grad, want)
        e_orc = rel_err(sd[n].grad, want)
        worst = max(worst, e_ours)
        assert e_ours <= 2.0 * e_orc + 3e-2, (n, e_ours, e_orc)
    assert worst < 0.15


def test_gradient_checkpointing_and_frozen_language(golden):
    need_gpu()
    g = golden("decoder_tiny")
    inp = {k: v.to(dev) for k, v in g["inputs"].items()}
    m1, m2 = build(g).train(), build(g).train()
    m2.gradient_checkpointing_enable()
    for m in (m1, m2):       # frozen_language: only parameters whose name contains "vision" train (modeling_libra.py:1342-1346)
        for n, p in m.named_parameters():
            p.requires_grad = "vision" in n
    kw = dict(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], vision_indices=inp["vision_indices"],
              contiguous_signal=inp["contiguous_signal"], labels=inp["labels"])
    l1, l2 = m1(**kw).loss, m2(**kw).loss
    assert float(l1) == float(l2)
    l1.backward(); l2.backward()
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        if "vision" in n and n != "vision_hidden_placeholder":      # the placeholder only feeds the disabled "2d" head
            assert p1.grad is not None and torch.equal(p1.grad, p2.grad), n      # recompute is bit-identical (deterministic kernels)
        elif "vision" not in n:
            assert p1.grad is None


def test_other_layouts_vs_oracle(golden):
    """two images per sample + left padding + explicit position ids, against the oracle in bf16 on the GPU."""
    need_gpu()
    from oracle.make_golden import make_libra_inputs
    g = golden("decoder_tiny")
    model = build(g).eval()
    cfg = g["config"]
    inp = make_libra_inputs(cfg["vocab_size"], cfg["contiguous_signal_size"], B=2, n_text=30, pad_last=11, seed=77, images_per_sample=2)
    with torch.no_grad():
        out = model(input_ids=inp["input_ids"].to(dev), attention_mask=inp["attention_mask"].to(dev),
                    vision_indices=inp["vision_indices"].to(dev), contiguous_signal=inp["contiguous_signal"].to(dev),
                    labels=inp["labels"].to(dev))
        orc = oracle_bf16(g, inp)
        sd32 = {k: v.to(dev) for k, v in g["state_dict"].items()}
        o32 = O.libra_forward(sd32, O.LibraDims.from_config(cfg), inp["input_ids"].to(dev), inp["vision_indices"].to(dev),
                              attention_mask=inp["attention_mask"].to(dev), contiguous_signal=inp["contiguous_signal"].to(dev),
                              labels=inp["labels"].to(dev))
    m = inp["attention_mask"].to(dev).bool()[None, :, :, None]
    fin = torch.isfinite(o32["logits"]) & m
    e_ours = rel_err(out.logits.float()[fin], o32["logits"][fin])
    e_orc = rel_err(orc["logits"].float()[fin], o32["logits"][fin])
    _record("vs_fp32_oracle", logits_ours=e_ours, logits_oracle_bf16=e_orc, loss_ours=abs(float(out.loss) - float(o32["loss"])))
    assert e_ours <= FACTOR * e_orc + SLACK, (e_ours, e_orc)
    assert abs(float(out.loss) - float(o32["loss"])) < 2e-3                                                  # measured 1.7e-5


def test_text_only_batch_forward_backward(golden):
    """No image in the batch (instruction-tuning samples without pictures): the vision segment of every routed op is empty.
    Loss and the gradients of the language weights against the bf16 oracle's autograd; the vision experts get no gradient."""
    need_gpu()
    g = golden("decoder_tiny")
    cfg = g["config"]
    V = cfg["vocab_size"]
    gen = torch.Generator().manual_seed(31)
    B, T = 2, 150
    ids = torch.randint(3, V, (B, T), generator=gen)
    ids[:, 0] = 1
    ids = ids[None].repeat(2, 1, 1).to(dev)
    am = torch.ones(B, T, dtype=torch.long)
    am[1, -13:] = 0
    am = am.to(dev)
    vi = torch.full((B, T), 578, device=dev)
    labels = ids.clone()
    labels[:, am == 0] = -100
    labels[labels == 1] = -100
    model = build(g).train()
    out = model(input_ids=ids, attention_mask=am, vision_indices=vi, labels=labels)
    out.loss.backward()
    sd = {k: (v.to(dev).bfloat16().requires_grad_(True) if v.is_floating_point() else v.to(dev)) for k, v in g["state_dict"].items()}
    o = O.libra_forward(sd, O.LibraDims.from_config(cfg), ids, vi, attention_mask=am, labels=labels)
    o["loss"].backward()
    assert abs(float(out.loss) - float(o["loss"])) < 5e-2
    params = dict(model.named_parameters())
    for name in ("model.layers.1.self_attn.q_proj.weight", "model.layers.0.mlp.down_proj.weight", "model.layers.1.input_layernorm.weight",
                 "lm_head.weight"):
        got, want = params[name].grad.float(), sd[name].grad.float()
        assert rel_err(got, want) < 8e-2, (name, rel_err(got, want))
    for name, p in params.items():
        if "vision" in name and p.grad is not None:
            assert float(p.grad.float().abs().max()) == 0.0, name


def test_mislabelled_targets_give_inf_like_the_reference(golden):
    need_gpu()
    g = golden("decoder_tiny")
    inp = {k: v.to(dev) for k, v in g["inputs"].items()}
    labels = g["inputs"]["input_ids"].clone().to(dev)      # raw next-token labels: BOI after text is a vision id on a language row
    model = build(g).train()
    out = model(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], vision_indices=inp["vision_indices"],
                contiguous_signal=inp["contiguous_signal"], labels=labels)
    assert torch.isinf(out.loss)


def test_left_padding_with_position_ids_vs_oracle(golden):
    """generation-style batch: sample 1 left padded, position_ids = cumsum(mask)-1 with pads at 1
    (libra/models/libra/modeling_libra.py:1204-1205); compared on the valid positions."""
    need_gpu()
    from oracle.make_golden import make_libra_inputs
    g = golden("decoder_tiny")
    cfg = g["config"]
    model = build(g).eval()
    inp = make_libra_inputs(cfg["vocab_size"], cfg["contiguous_signal_size"], B=2, n_text=20, pad_last=0, seed=9)
    am = inp["attention_mask"].clone()
    ids, vi = inp["input_ids"].clone(), inp["vision_indices"].clone()
    npad = 5
    # shift sample 1 right by npad: pads (id 0, language) in front
    ids[:, 1] = torch.cat([torch.zeros(2, npad, dtype=ids.dtype), ids[:, 1, :-npad]], dim=1)
    vi[1] = torch.cat([torch.full((npad,), 578), vi[1, :-npad]])
    sig = inp["contiguous_signal"].clone()
    sig[1] = torch.cat([torch.zeros(npad, sig.shape[-1]), sig[1, :-npad]], dim=0)
    am[1, :npad] = 0
    # the shifted image lost its tail: keep flags consistent (truncated image tokens are still vision ids)
    assert torch.equal(vi < 578, ids[0] >= cfg["vocab_size"])
    pos = am.long().cumsum(-1) - 1
    pos.masked_fill_(am == 0, 1)
    with torch.no_grad():
        out = model(input_ids=ids.to(dev), attention_mask=am.to(dev), vision_indices=vi.to(dev), position_ids=pos.to(dev),
                    contiguous_signal=sig.to(dev))
        sd32 = {k: v.to(dev) for k, v in g["state_dict"].items()}
        o32 = O.libra_forward(sd32, O.LibraDims.from_config(cfg), ids.to(dev), vi.to(dev), attention_mask=am.to(dev),
                              contiguous_signal=sig.to(dev), position_ids=pos.to(dev))
        sd16 = {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in sd32.items()}
        o16 = O.libra_forward(sd16, O.LibraDims.from_config(cfg), ids.to(dev), vi.to(dev), attention_mask=am.to(dev),
                              contiguous_signal=sig.to(dev).bfloat16(), position_ids=pos.to(dev))
    m = am.to(dev).bool()[None, :, :, None]
    fin = torch.isfinite(o32["logits"]) & m
    e_ours = rel_err(out.logits.float()[fin], o32["logits"][fin])
    e_orc = rel_err(o16["logits"].float()[fin], o32["logits"][fin])
    _record("left_padding", logits_ours=e_ours, logits_oracle_bf16=e_orc)
    assert e_ours <= FACTOR * e_orc + SLACK, (e_ours, e_orc)


def test_full_width_layer_vs_oracle_fp32_and_bf16():
    """One full-width Libra-11B layer inside LibraForCausalLM (H 4096, 32 heads, I 11008, 32514 logits), T = 2048, one image
    + right padding, forward AND backward, against the oracle in fp32 with the bf16 oracle as the noise floor
    (scripts/parity_report.py; measured numbers of the same run are committed in profiles/r02_parity.json).
    Asserted: the -inf pattern is the reference's; ours is no further from fp32 than 1.25x the bf16 reference + 1e-3 on the
    logits (rel. Frobenius), the loss within 5e-3 absolute, every gradient within 1.5x the bf16 oracle + 1e-2."""
    need_gpu()
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import parity_report
    r = parity_report.report(2048)
    lg = r["logits"]
    assert lg["inf_pattern_matches"]
    assert lg["rel_fro_ours_vs_fp32"] <= 1.25 * lg["rel_fro_oracle_bf16_vs_fp32"] + 1e-3, lg
    assert lg["frac_within_1e-3_rel_ours"] >= lg["frac_within_1e-3_rel_oracle_bf16"] - 0.02, lg
    assert r["loss"]["abs_err_ours"] <= max(2.0 * r["loss"]["abs_err_oracle_bf16"], 5e-3), r["loss"]
    for n, v in r["grads"].items():
        assert v["rel_fro_ours_vs_fp32"] <= 1.5 * v["rel_fro_oracle_bf16_vs_fp32"] + 1e-2, (n, v)
