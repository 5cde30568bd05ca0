"""Host-side scheduling of the persistent attention forward (libra_b200/schedule.py:stream_plan): the (work item, head) list
must be split so that every item runs exactly once, in head-group order, with balanced per-CTA loads and a correct
`longest per-CTA list` (the kernel's item table is sized by it)."""
import torch

from libra_b200 import schedule


def _work(B, T, spans, kv_end=None):
    flag = torch.zeros(B, T, dtype=torch.bool)
    for b, s, e in spans:
        flag[b, s:e] = True
    return schedule.build_attn_work(flag, B, T, True, "cpu", kv_end=kv_end)


def test_stream_plan_is_a_balanced_partition():
    w = _work(4, 2048, [(b, 1, 579) for b in range(4)])
    heads, n_cta, group = 32, 148, 8
    items, off, n, longest = w.stream_plan(heads, n_cta, group)
    n_items = len(w.q_tiles) * heads
    assert n == n_cta and off.shape[0] == n_cta + 1 and int(off[0]) == 0 and int(off[-1]) == n_items
    assert sorted(items.tolist()) == list(range(n_items)), "every (work item, head) exactly once"
    sizes = (off[1:] - off[:-1]).tolist()
    assert max(sizes) == longest
    # load = kv tiles (+ per-item overhead) per CTA: within a few tiles of each other
    per_group = group * len(w.q_tiles)
    def tiles_of(L):
        rem = L % per_group
        return w.q_tiles[rem // min(group, heads)]
    loads = [sum(tiles_of(L) + 2.0 for L in items[int(off[c]):int(off[c + 1])].tolist()) for c in range(n_cta)]
    assert max(loads) - min(loads) <= 16 + 2.0, (min(loads), max(loads))
    # head groups are dealt in order: a CTA's list never goes back to an earlier group
    for c in range(n_cta):
        g = [L // per_group for L in items[int(off[c]):int(off[c + 1])].tolist()]
        assert g == sorted(g)
    # cached per (heads, n_cta, group)
    assert w.stream_plan(heads, n_cta, group)[0] is items


def test_stream_plan_small_lists_and_ragged_batches():
    w = _work(2, 700, [(0, 1, 579), (1, 50, 628)], kv_end=[667, 700])
    items, off, n, longest = w.stream_plan(2, 148, 8)
    n_items = len(w.q_tiles) * 2
    assert n == min(148, n_items) and int(off[-1]) == n_items and longest == 1
    assert sorted(items.tolist()) == list(range(n_items))
    # mixed-modality q tiles appear once per variant; every (sample, q tile) that holds rows is covered
    wq = w.work_q.tolist()
    assert len({(b, qt, v) for b, qt, v, _ in wq}) == len(wq)
    assert {(b, qt) for b, qt, _, _ in wq} == {(b, qt) for b in range(2) for qt in range((700 + 127) // 128)}
