"""N1: KV-cached decoding on the CUDA path.

* lb_attn_decode vs a plain PyTorch fp32 restatement of the same one-query attention over random caches (ragged key ranges,
  both variants, several splits);
* LibraForCausalLM(use_cache=True): prefill + one-token steps vs the fixture the REFERENCE produced with its own cache tuple
  (tests/golden/decode_tiny.pt, oracle/make_golden.py:golden_decode) and vs the oracle's libra_forward_cached in bf16;
* stepwise decoding == full-sequence forward (the cache must not change the function);
* greedy generate() follows the reference's vision-index bookkeeping.
Tolerances as in test_gpu_model.py: bf16 model against the fp32 reference, the bf16 oracle's own distance is the noise floor."""
import math

import pytest
import torch

from gpu_util import need_gpu, rel_err, assert_close
from oracle import libra_oracle as O

pytestmark = pytest.mark.gpu
dev = "cuda"


@pytest.mark.parametrize("B,H,D,T,cap", [(2, 2, 128, 300, 512), (3, 4, 128, 1500, 1536), (2, 3, 64, 77, 128), (1, 1, 128, 1, 8)])
def test_attn_decode_kernel(B, H, D, T, cap):
    need_gpu()
    from libra_b200 import ops
    g = torch.Generator(device=dev).manual_seed(B * 1000 + T)
    C = H * D
    mk = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
    q = mk(B, C)
    Kfl, Vfl, Kfv, Vfv = mk(B, cap, C), mk(B, cap, C), mk(B, cap, C), mk(B, cap, C)
    qflag = (torch.arange(B, device=dev) % 2).to(torch.uint8)
    kv_start = torch.tensor([(3 * b) % max(T // 2, 1) for b in range(B)], device=dev, dtype=torch.int32)
    kv_end = torch.tensor([T - (b % 2) * min(5, T - 1) for b in range(B)], device=dev, dtype=torch.int32)
    out_row = torch.arange(B - 1, -1, -1, device=dev, dtype=torch.int32)
    scale = 1 / math.sqrt(D)
    out = ops.attn_decode(q, Kfl, Vfl, Kfv, Vfv, qflag, kv_start, kv_end, out_row, B, H, D, T, scale)
    torch.cuda.synchronize()
    for b in range(B):
        K, V = (Kfv, Vfv) if int(qflag[b]) else (Kfl, Vfl)
        s, e = int(kv_start[b]), int(kv_end[b])
        k = K[b, s:e].float().view(e - s, H, D)
        v = V[b, s:e].float().view(e - s, H, D)
        sc = torch.einsum("hd,thd->ht", q[b].float().view(H, D), k) * scale
        p = torch.softmax(sc, dim=-1).bfloat16().float()
        want = torch.einsum("ht,thd->hd", p, v).reshape(C)
        assert_close(out[int(out_row[b])], want, rtol=2e-2, atol=2e-2, msg=f"sample {b}")


def _build(g):
    from libra_b200.models import LibraConfig, LibraForCausalLM
    model = LibraForCausalLM(LibraConfig(**g["config"]))
    model.load_state_dict(g["state_dict"], strict=False)
    return model.to(torch.bfloat16).to(dev).eval()


def _run_steps(model, gd):
    inp = gd["inputs"]
    am = inp["attention_mask"].to(dev)
    out = model(input_ids=inp["input_ids"].to(dev), attention_mask=am, vision_indices=inp["vision_indices"].to(dev),
                contiguous_signal=inp["contiguous_signal"].to(dev), use_cache=True)
    logits = [out.logits[:, :, -1].float()]
    for nid, nvi in zip(gd["step_input_ids"], gd["step_vision_indices"]):
        am = torch.cat([am, am.new_ones(am.shape[0], 1)], dim=1)
        pos = (am.long().cumsum(-1) - 1)[:, -1:]
        out = model(input_ids=nid.to(dev), attention_mask=am, vision_indices=nvi.to(dev), position_ids=pos,
                    past_key_values=out.past_key_values, use_cache=True)
        assert out.logits.shape[2] == 1
        logits.append(out.logits[:, :, -1].float())
    return torch.stack(logits), out.past_key_values


def test_cached_decode_vs_reference_golden(golden):
    need_gpu()
    gm, gd = golden("decoder_tiny"), golden("decode_tiny")
    model = _build(gm)
    got, cache = _run_steps(model, gd)
    want = gd["logits"].to(dev)
    fin = torch.isfinite(want)
    assert torch.equal(fin, torch.isfinite(got)), "the -inf / newline placeholder pattern must match the reference"
    # the bf16 oracle, same steps, is the noise floor
    sd = {k: (v.to(dev).bfloat16() if v.is_floating_point() else v.to(dev)) for k, v in gm["state_dict"].items()}
    d = O.LibraDims.from_config(gm["config"])
    inp = gd["inputs"]
    am = inp["attention_mask"].to(dev)
    o = O.libra_forward_cached(sd, d, inp["input_ids"].to(dev), inp["vision_indices"].to(dev), attention_mask=am,
                               contiguous_signal=inp["contiguous_signal"].to(dev).bfloat16())
    orc = [o["logits"][:, :, -1].float()]
    for nid, nvi in zip(gd["step_input_ids"], gd["step_vision_indices"]):
        am = torch.cat([am, am.new_ones(am.shape[0], 1)], dim=1)
        pos = (am.long().cumsum(-1) - 1)[:, -1:]
        o = O.libra_forward_cached(sd, d, nid.to(dev), nvi.to(dev), attention_mask=am, position_ids=pos,
                                   past_key_values=o["past_key_values"])
        orc.append(o["logits"][:, :, -1].float())
    orc = torch.stack(orc)
    for s in range(want.shape[0]):
        f = fin[s]
        e_ours, e_orc = rel_err(got[s][f], want[s][f]), rel_err(orc[s][f], want[s][f])
        assert e_ours <= 1.5 * e_orc + 5e-3, (s, e_ours, e_orc)
    # </img> just consumed => "append a newline": +inf on the newline token only
    assert torch.isposinf(got[4, :, 1, model.config.newline_token_id]).all()
    # the cache in the reference's layout: last 8 positions of layer 1
    ref_c = cache.to_reference()
    assert rel_err(ref_c[1][0][1][:, :, -8:].float(), gd["k_for_language_l1"].to(dev)) < 2e-2
    assert rel_err(ref_c[1][0][0][:, :, -8:].float(), gd["k_for_vision_l1"].to(dev)) < 2e-2
    assert cache.get_seq_length() == inp["input_ids"].shape[2] + len(gd["step_input_ids"])


def test_stepwise_equals_full_forward(golden):
    """Decoding token by token must give the logits the training-path forward gives on the whole sequence."""
    need_gpu()
    gm, gd = golden("decoder_tiny"), golden("decode_tiny")
    model = _build(gm)
    got, _ = _run_steps(model, gd)
    inp = gd["inputs"]
    ids = torch.cat([inp["input_ids"]] + [t for t in gd["step_input_ids"]], dim=2).to(dev)
    vi = torch.cat([inp["vision_indices"]] + [t for t in gd["step_vision_indices"]], dim=1).to(dev)
    T0 = inp["input_ids"].shape[2]
    sig = inp["contiguous_signal"]
    sig = torch.cat([sig, sig.new_zeros(sig.shape[0], len(gd["step_input_ids"]), sig.shape[2])], dim=1).to(dev)   # decode steps carry no signal (:1216-1218)
    with torch.no_grad():
        full = model(input_ids=ids, vision_indices=vi, contiguous_signal=sig, use_cache=False).logits.float()
    for s in range(got.shape[0] - 2):            # (the </img> step and the one after it differ by the newline placeholder)
        a, b = got[s], full[:, :, T0 - 1 + s]
        f = torch.isfinite(b)
        assert torch.equal(f, torch.isfinite(a))
        assert rel_err(a[f], b[f]) < 2e-2, s


def test_left_padded_batch_and_greedy_generate(golden):
    need_gpu()
    gm = golden("decoder_tiny")
    model = _build(gm)
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(9)
    T = 24
    ids = torch.randint(3, V, (2, T), generator=g)
    am = torch.ones(2, T, dtype=torch.long)
    am[1, :7] = 0
    ids[1, :7] = 0
    ids = ids[None].repeat(2, 1, 1).to(dev)
    vi = torch.full((2, T), 578, device=dev)
    out = model.generate(ids, attention_mask=am.to(dev), vision_indices=vi, max_new_tokens=6)       # CUDA-graph replay of the step
    assert out.shape == (2, 2, T + 6)
    eager = model.generate(ids, attention_mask=am.to(dev), vision_indices=vi, max_new_tokens=6, cuda_graph=False)
    assert torch.equal(out, eager), "the captured step must reproduce the eager one-token steps bit for bit"
    longer = model.generate(ids, attention_mask=am.to(dev), vision_indices=vi, max_new_tokens=300)   # forces the cache to grow
    assert torch.equal(longer[:, :, :T + 6], out)
    # EOS bookkeeping inside the graph: pick a token the greedy path emits for sample 0 and call it EOS
    eos = int(out[0, 0, T + 2])
    a = model.generate(ids, attention_mask=am.to(dev), vision_indices=vi, max_new_tokens=40, eos_token_id=eos)
    b = model.generate(ids, attention_mask=am.to(dev), vision_indices=vi, max_new_tokens=40, eos_token_id=eos, cuda_graph=False)
    n = min(a.shape[2], b.shape[2])
    assert torch.equal(a[:, :, :n], b[:, :, :n])
    first = (a[0, 0, T:] == eos).nonzero()[0, 0]
    assert (a[:, 0, T + first:] == eos).all()          # a finished sample keeps emitting EOS
    # the same continuation without the cache: greedy on the full forward, one token at a time
    cur, cam = ids, am.to(dev)
    for _ in range(6):
        pos = cam.cumsum(-1) - 1
        pos.masked_fill_(cam == 0, 1)
        cvi = torch.full_like(cur[0], 578)          # language rows can only predict text ids (the vision block is -inf)
        lg = model(input_ids=cur, attention_mask=cam, position_ids=pos, vision_indices=cvi, use_cache=False).logits[:, :, -1].float()
        nxt = lg.argmax(-1)
        cur = torch.cat([cur, nxt[:, :, None]], dim=2)
        cam = torch.cat([cam, cam.new_ones(2, 1)], dim=1)
    assert (out < V).all()
    assert (out == cur).float().mean() > 0.95       # bf16 ties aside, cached and uncached greedy paths agree


def test_decode_with_and_without_dependent_launch(golden, monkeypatch):
    """The one-token steps run as a programmatic-dependent-launch chain (lb_set_pdl): LB_PDL=0 (serial launches) and the
    default give the same tokens, eagerly and through the CUDA graph."""
    need_gpu()
    gm = golden("decoder_tiny")
    model = _build(gm)
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(4)
    T = 33
    ids = torch.randint(3, V, (3, T), generator=g)[None].repeat(2, 1, 1).to(dev)
    vi = torch.full((3, T), 578, device=dev)
    outs = {}
    for pdl in ("1", "0"):
        monkeypatch.setenv("LB_PDL", pdl)
        for cg in (True, False):
            outs[(pdl, cg)] = model.generate(ids, vision_indices=vi, max_new_tokens=24, cuda_graph=cg)
    ref = outs[("0", False)]
    for k, v in outs.items():
        assert torch.equal(v, ref), k


def test_sampling_processors_and_policy_in_generate(golden):
    """generate(): the reference's `sample` loop on this path (modeling_libra_utils.py:330-635) -- built-in warpers, caller
    processors, repetition penalty inside the captured step, plane-ordered pad/EOS rule, argument checking."""
    need_gpu()
    gm = golden("decoder_tiny")
    model = _build(gm)
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(21)
    T = 19
    ids = torch.randint(3, V, (2, T), generator=g)[None].repeat(2, 1, 1).to(dev)
    vi = torch.full((2, T), 578, device=dev)
    kw = dict(vision_indices=vi, max_new_tokens=12)
    greedy = model.generate(ids, **kw)
    # top_k = 1 sampling picks a maximiser of the processed scores at every step (bf16 logits tie, so the SEQUENCE may differ
    # from argmax's first-index choice); through the CUDA graph (multinomial captured) a fixed seed reproduces the run
    o = model.generate(ids, do_sample=True, top_k=1, return_dict_in_generate=True, output_scores=True, **kw)
    for t, sc in enumerate(o.scores):
        tok = o.sequences[:, :, T + t]
        assert torch.equal(sc.gather(2, tok[:, :, None])[..., 0], sc.max(dim=-1).values), t
        assert (torch.isfinite(sc).sum(-1) >= 1).all()
    torch.manual_seed(5)
    g1 = model.generate(ids, do_sample=True, temperature=0.9, **kw)
    torch.manual_seed(5)
    g2 = model.generate(ids, do_sample=True, temperature=0.9, **kw)
    assert torch.equal(g1, g2)
    # repetition penalty: the captured step (fixed-size history buffer) == the eager loop
    a = model.generate(ids, repetition_penalty=1.7, **kw)
    b = model.generate(ids, repetition_penalty=1.7, cuda_graph=False, **kw)
    assert torch.equal(a, b)
    # a seeded generator reproduces the eager sampling loop; tokens are valid text ids
    s1 = model.generate(ids, do_sample=True, temperature=0.8, top_p=0.9, generator=torch.Generator(dev).manual_seed(3), **kw)
    s2 = model.generate(ids, do_sample=True, temperature=0.8, top_p=0.9, generator=torch.Generator(dev).manual_seed(3), **kw)
    assert torch.equal(s1, s2) and s1.shape == greedy.shape and int(s1[:, :, T:].max()) < V and int(s1.min()) >= 0
    s3 = model.generate(ids, do_sample=True, temperature=1.5, **kw)          # default generator, graph path
    assert s3.shape == greedy.shape and int(s3[:, :, T:].max()) < V
    # a caller-supplied processor (applied per plane on [B, V] scores) forces the outcome
    def only_seven(input_ids, scores):
        assert input_ids.dim() == 2 and scores.dim() == 2
        out = torch.full_like(scores, -float("inf"))
        out[:, 7] = 0.0
        return out
    forced = model.generate(ids, logits_processor=[only_seven], **kw)
    assert (forced[:, :, T:] == 7).all()
    # EOS / pad: a finished sample is padded with pad_token_id, not EOS, when one is given
    eos = int(greedy[0, 0, T + 3])
    p = model.generate(ids, eos_token_id=eos, pad_token_id=0, max_new_tokens=12, vision_indices=vi, cuda_graph=False)
    first = int((p[0, 0, T:] == eos).nonzero()[0, 0])
    assert (p[:, 0, T + first + 1:] == 0).all()
    out = model.generate(ids, return_dict_in_generate=True, output_scores=True, **kw)
    assert torch.equal(out.sequences, greedy) and len(out.scores) == 12 and out.scores[0].shape == (2, 2, V + 514)
    with pytest.raises(TypeError):
        model.generate(ids, penalty_alpha=0.6, **kw)
    with pytest.raises(NotImplementedError):
        model.generate(ids, num_beams=4, **kw)


def test_decode_with_the_bridge_folded_into_the_prologue(golden, monkeypatch):
    """LB_FOLD_DECODE_BRIDGE: the rank-8 bridge products computed inside lb_attn_prep_fwd_bridge instead of a skinny-GEMM launch;
    same rounding sequence, fp32 summation order aside -- the greedy continuation agrees (ties in bf16 logits aside)."""
    need_gpu()
    from libra_b200 import functional as LF
    gm = golden("decoder_tiny")
    model = _build(gm)
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(6)
    T = 21
    ids = torch.randint(3, V, (2, T), generator=g)[None].repeat(2, 1, 1).to(dev)
    vi = torch.full((2, T), 578, device=dev)
    ref = model.generate(ids, vision_indices=vi, max_new_tokens=16)
    monkeypatch.setattr(LF, "FOLD_DECODE_BRIDGE", True)
    for cg in (True, False):
        out = model.generate(ids, vision_indices=vi, max_new_tokens=16, cuda_graph=cg)
        assert out.shape == ref.shape and (out == ref).float().mean() > 0.9
