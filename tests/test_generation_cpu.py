"""N1 host logic: libra_b200/generation.py against transformers' own logits processors / warpers (the classes the reference's
greedy_search / sample loops receive, modeling_libra_utils.py:61-635) and against the reference loop's finished-sample rule."""
import pytest
import torch

from libra_b200 import generation as G

tfm = pytest.importorskip("transformers")


def _data(seed=0, Q=2, B=3, T=11, V=97):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, V, (Q, B, T), generator=g)
    logits = torch.randn(Q, B, V, generator=g) * 3
    return ids, logits


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_builtin_policy_equals_transformers_classes(seed):
    from transformers import (RepetitionPenaltyLogitsProcessor, TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper)
    ids, logits = _data(seed)
    pol = G.SelectionPolicy(do_sample=True, temperature=0.7, top_k=20, top_p=0.9, repetition_penalty=1.3)
    got = G.process(pol, ids, logits)
    chain = [RepetitionPenaltyLogitsProcessor(1.3), TemperatureLogitsWarper(0.7), TopKLogitsWarper(20), TopPLogitsWarper(0.9)]
    for q in range(ids.shape[0]):
        s = logits[q].float()
        for c in chain:
            s = c(ids[q], s)
        assert torch.equal(got[q], s), q
    # greedy: warpers are not applied (greedy_search only runs logits_processor), the penalty is
    pol = G.SelectionPolicy(do_sample=False, temperature=0.7, top_k=20, top_p=0.9, repetition_penalty=1.3)
    got = G.process(pol, ids, logits)
    for q in range(ids.shape[0]):
        assert torch.equal(got[q], RepetitionPenaltyLogitsProcessor(1.3)(ids[q], logits[q].float()))
    assert torch.equal(G.select(pol, got), got.argmax(-1))


def test_user_processor_lists_run_per_plane():
    from transformers import LogitsProcessorList, MinLengthLogitsProcessor, TopKLogitsWarper
    ids, logits = _data(5)
    procs = LogitsProcessorList([MinLengthLogitsProcessor(50, eos_token_id=2)])
    pol = G.SelectionPolicy(do_sample=True, logits_processor=procs, logits_warper=[TopKLogitsWarper(5)])
    got = G.process(pol, ids, logits)
    assert torch.isinf(got[:, :, 2]).all()                          # EOS suppressed before min_length
    assert ((~torch.isinf(got)).sum(-1) == 5).all()
    tok = G.select(pol, got)
    assert tok.shape == (2, 3) and (torch.gather(got, 2, tok[:, :, None]) > -float("inf")).all()


def test_sampling_follows_the_reference_draw_order():
    """One torch.multinomial per plane, plane 0 first (modeling_libra_utils.py:559-564): same generator state => same tokens."""
    ids, logits = _data(7)
    pol = G.SelectionPolicy(do_sample=True, temperature=1.3, generator=torch.Generator().manual_seed(123))
    got = G.next_tokens(pol, ids, logits)
    g = torch.Generator().manual_seed(123)
    probs = torch.softmax(logits.float() / 1.3, dim=-1)
    want = torch.stack([torch.multinomial(p, num_samples=1, generator=g).squeeze(1) for p in probs])
    assert torch.equal(got, want)
    # empirical distribution of one row
    pol = G.SelectionPolicy(do_sample=True, top_k=3, generator=torch.Generator().manual_seed(1))
    row = torch.tensor([[[2.0, 1.0, 0.0, -1.0, -2.0]]]).repeat(1, 4000, 1)
    tok = G.next_tokens(pol, torch.zeros(1, 4000, 1, dtype=torch.long), row)
    freq = torch.bincount(tok[0], minlength=5).float() / 4000
    want = torch.softmax(torch.tensor([2.0, 1.0, 0.0]), 0)
    assert freq[3:].sum() == 0 and (freq[:3] - want).abs().max() < 0.03


def test_finished_samples_follow_the_reference_loop():
    """Reference (modeling_libra_utils.py:266-287): per plane in order, token = token*unfinished + pad*(1-unfinished), then
    unfinished *= (token != eos)."""
    eos, pad = 2, 0
    nxt = torch.tensor([[5, 2, 7, 2], [6, 9, 2, 2]])
    done = torch.tensor([False, False, False, True])
    want, unf = [], (~done).long()
    for row in nxt:
        r = row * unf + pad * (1 - unf)
        unf = unf * (r != eos).long()
        want.append(r)
    got, d = G.finish_(nxt.clone(), done.clone(), torch.tensor([eos]), pad)
    assert torch.equal(got, torch.stack(want)) and torch.equal(d, unf == 0)
    same, d2 = G.finish_(nxt.clone(), done.clone(), None, None)
    assert torch.equal(same, nxt) and torch.equal(d2, done)


def test_policy_validation():
    for kw in (dict(temperature=0.0), dict(top_p=1.5), dict(top_k=-1), dict(repetition_penalty=0.0)):
        with pytest.raises(ValueError):
            G.SelectionPolicy(**kw)
