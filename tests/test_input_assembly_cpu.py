"""N4 (first half): the sync-free tensor assembly and label construction against the oracle's restatement of the reference
(tokenization_libra.py:250-316, modeling_libra.py:1397-1411) -- integer results bit exact, signal rows exact."""
import pytest
import torch

from oracle import libra_oracle as O


def _case(seed, B=3, n_img=(1, 0, 2), n_text=20, L=10, S=6):
    g = torch.Generator().manual_seed(seed)
    T = 1 + max(n_img) * L + n_text
    text = torch.randint(3, 300, (B, T), generator=g)
    for b, k in enumerate(n_img):
        pos = 1 + b
        for _ in range(k):
            text[b, pos:pos + L] = 999
            pos += L + 2
    n = sum(n_img)
    image_ids = torch.randint(320, 832, (2, n, L), generator=g)
    feat = torch.randn(n, L - 2, S, generator=g)
    am = torch.ones(B, T, dtype=torch.long)
    am[-1, -4:] = 0
    return text, am, image_ids, feat, L


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("ignore", [False, True])
def test_assemble_inputs_matches_oracle(seed, ignore):
    from libra_b200.models.tokenization_libra import assemble_inputs
    text, am, image_ids, feat, L = _case(seed)
    ign = torch.tensor([True, False, True]) if ignore else None
    got = assemble_inputs(text, am, 999, image_ids, feat, max_vision_token_length=L, contiguous_ignore=ign, check=True)
    want = O.assemble_inputs(text, am, 999, image_ids, feat, max_vision_token_length=L, contiguous_ignore=ign)
    for k in ("input_ids", "attention_mask", "vision_indices", "coninous_signal"):
        assert torch.equal(got[k], want[k]), k
    short = assemble_inputs(text, am, 999, image_ids, feat, max_vision_token_length=L, truncation=True, max_length=17)
    assert short["input_ids"].shape[2] == 17 and torch.equal(short["vision_indices"], want["vision_indices"][:, :17])
    with pytest.raises(ValueError):
        assemble_inputs(text, am, 999, image_ids[:, :-1], feat[:-1], max_vision_token_length=L, check=True)


def test_text_only_and_get_labels():
    from libra_b200.models.tokenization_libra import assemble_inputs, get_labels
    text, am, _, _, L = _case(5, n_img=(0, 0, 0))
    out = assemble_inputs(text, am, 999, None, None, max_vision_token_length=L)
    assert out["coninous_signal"] is None and (out["vision_indices"] == L).all() and torch.equal(out["input_ids"][1], text)
    ids = torch.randint(3, 300, (2, 2, 12))
    ids[:, :, 0] = 1
    ids[:, 0, 3] = 832
    am2 = torch.ones(2, 12, dtype=torch.long)
    am2[1, -2:] = 0
    spans = [[[4, 6]], [[1, 2], [7, 9]]]
    assert torch.equal(get_labels(ids, am2, 832, 1, spans), O.get_labels(ids, am2, 832, 1, spans))
