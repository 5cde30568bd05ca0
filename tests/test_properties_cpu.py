"""Property tests (hypothesis) of the host logic the kernels rely on, and of the preprocessing oracle against Pillow on random
sizes.  CPU only; small example counts keep the file under ~20 s."""
import numpy as np
import pytest
import torch

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from libra_b200 import schedule  # noqa: E402
from oracle import clip_preprocess_oracle as O  # noqa: E402

TILE = 128


@st.composite
def layouts(draw):
    """A batch layout like the reference's tokenizer produces: per sample some image spans (vision rows) inside text."""
    B = draw(st.integers(1, 3))
    T = draw(st.integers(1, 700))
    flag = torch.zeros(B, T, dtype=torch.bool)
    for b in range(B):
        for _ in range(draw(st.integers(0, 3))):
            s = draw(st.integers(0, T - 1))
            e = draw(st.integers(s, min(T, s + 300)))
            flag[b, s:e] = True
    return flag


@settings(max_examples=40, deadline=None)
@given(layouts())
def test_routing_is_a_stable_partition(flag):
    """build_routing (the sorted-row layout that replaces cal_language_vision's boolean masks, modeling_libra.py:111-147):
    language rows first, vision rows second, original order kept inside each segment, inv o perm = identity."""
    rt = schedule.build_routing(flag)
    f = flag.reshape(-1)
    n = f.numel()
    perm, inv = rt.perm.long(), rt.inv.long()
    assert rt.n_tokens == n and rt.n_vis == int(f.sum()) and rt.n_lang == n - rt.n_vis
    assert sorted(perm.tolist()) == list(range(n))
    assert torch.equal(inv[perm], torch.arange(n)) and torch.equal(perm[inv], torch.arange(n))
    assert not f[perm[:rt.n_lang]].any() and f[perm[rt.n_lang:]].all()
    assert torch.equal(perm[:rt.n_lang], torch.sort(perm[:rt.n_lang]).values)          # stable
    assert torch.equal(perm[rt.n_lang:], torch.sort(perm[rt.n_lang:]).values)
    assert torch.equal(rt.flag_sorted.bool(), f[perm]) and torch.equal(rt.flag_orig.bool(), f)


@settings(max_examples=40, deadline=None)
@given(layouts(), st.booleans())
def test_attention_work_lists_cover_exactly_what_the_mask_allows(flag, causal):
    """build_attn_work: a forward / dQ item exists for (sample, q tile, variant) iff the tile holds a row of that modality; a
    dK/dV item for (sample, kv tile, variant) iff some q tile at or after it (causal) holds such a row; the kv tile count of a
    forward item is the causal triangle's; kv_cover tells where dK/dV needs no zero fill."""
    B, T = flag.shape
    nt = (T + TILE - 1) // TILE
    w = schedule.build_attn_work(flag, B, T, causal, "cpu")
    has = torch.zeros(B, 2, nt, dtype=torch.bool)
    for b in range(B):
        for qt in range(nt):
            seg = flag[b, qt * TILE:(qt + 1) * TILE]
            has[b, 0, qt] = bool((~seg).any())
            has[b, 1, qt] = bool(seg.any())
    assert torch.equal(w.qtile_has.bool(), has)
    want_q = {(b, qt, v) for b in range(B) for v in range(2) for qt in range(nt) if has[b, v, qt]}
    got_q = [(b, qt, v) for b, qt, v, _ in w.work_q.tolist()]
    assert len(got_q) == len(set(got_q)) and set(got_q) == want_q
    for (b, qt, v, _), n_kv in zip(w.work_q.tolist(), w.q_tiles):
        assert n_kv == (qt + 1 if causal else nt)
    assert w.q_tiles == sorted(w.q_tiles, reverse=True)                                 # longest items first
    want_kv = {(b, kt, v) for b in range(B) for v in range(2) for kt in range(nt) if has[b, v, (kt if causal else 0):].any()}
    got_kv = [(b, kt, v) for b, kt, v, _ in w.work_kv.tolist()]
    assert len(got_kv) == len(set(got_kv)) and set(got_kv) == want_kv
    for v in range(2):
        full = all((b, kt, v) in want_kv for b in range(B) for kt in range(nt))
        assert w.kv_cover[v] == full


@settings(max_examples=25, deadline=None)
@given(layouts(), st.integers(1, 8), st.integers(1, 200))
def test_stream_plan_partitions_every_item_once(flag, heads, n_cta):
    B, T = flag.shape
    w = schedule.build_attn_work(flag, B, T, True, "cpu")
    for which, n_work in (("q", len(w.q_tiles)), ("kv", len(w.kv_tiles))):
        items, off, n, longest = w.stream_plan(heads, n_cta, 8, which=which)
        total = n_work * heads
        assert 1 <= n <= n_cta and off.shape[0] == n + 1 and int(off[0]) == 0 and int(off[-1]) == total
        assert sorted(items.tolist()) == list(range(total))
        sizes = (off[1:] - off[:-1]).tolist()
        assert min(sizes) >= 0 and max(sizes) == longest


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 260), st.integers(1, 260), st.sampled_from([24, 56, 57]), st.integers(0, 2 ** 31 - 1))
def test_preprocessing_oracle_equals_pillow_on_random_sizes(h, w, size, seed):
    """Pillow's BICUBIC through oracle/clip_preprocess_oracle.py at arbitrary (also degenerate) sizes: bit exact."""
    from PIL import Image
    oh, ow = O.resize_output_size(h, w, size)
    if oh * ow > 300_000:                      # extreme aspect ratios explode the long edge: keep the example cheap
        return
    img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
    assert np.array_equal(O.pil_resize_bicubic(img, ow, oh), ref)
    # the library's host tables for this axis pair are the oracle's
    import ctypes
    from libra_b200 import _lib
    for in_size, out_size in ((w, ow), (h, oh)):
        bounds, kk, ksize = O.precompute_coeffs(in_size, out_size)
        n = min(out_size, 64)
        cb, bb = (ctypes.c_int32 * (n * ksize))(), (ctypes.c_int32 * (2 * n))()
        assert _lib.load().lb_clip_resample_coeffs(in_size, out_size, 0, n, cb, n * ksize, bb) == ksize
        assert np.array_equal(np.frombuffer(cb, dtype=np.int32).reshape(n, ksize), kk[:n])
        assert np.array_equal(np.frombuffer(bb, dtype=np.int32).reshape(n, 2), bounds[:n])


@st.composite
def text_batches(draw):
    """Text ids with placeholder runs of length L per image (what the reference's text tokenizer hands to LibraTokenizer.forward,
    tokenization_libra.py:250-316), right padding on some samples."""
    B = draw(st.integers(1, 4))
    L = draw(st.integers(3, 12))
    n_img = [draw(st.integers(0, 3)) for _ in range(B)]
    gaps = [[draw(st.integers(0, 5)) for _ in range(k + 1)] for k in n_img]
    T = max(1 + sum(g) + k * L for g, k in zip(gaps, n_img)) + draw(st.integers(0, 6))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(3, 300, (B, T), generator=g)
    am = torch.ones(B, T, dtype=torch.long)
    for b in range(B):
        pos = 1
        for i in range(n_img[b]):
            pos += gaps[b][i]
            text[b, pos:pos + L] = 999
            pos += L
        used = pos + gaps[b][-1]
        if used < T and draw(st.booleans()):
            am[b, used:] = 0
    n = sum(n_img)
    image_ids = torch.randint(320, 832, (2, n, L), generator=g)
    feat = torch.randn(n, L - 2, 5, generator=g)
    return text, am, image_ids, feat, L, n


@settings(max_examples=40, deadline=None)
@given(text_batches())
def test_assemble_inputs_equals_oracle_on_random_layouts(case):
    """The sync-free assembly (rank-of-placeholder gathers, no nonzero) against the oracle's restatement of the reference's
    boolean-mask scatters: ids, mask, vision indices bit exact, signal rows exact."""
    from libra_b200.models.tokenization_libra import assemble_inputs
    from oracle import libra_oracle as LO
    text, am, image_ids, feat, L, n = case
    if n == 0:
        got = assemble_inputs(text, am, 999, None, None, max_vision_token_length=L)
        assert got["coninous_signal"] is None and (got["vision_indices"] == L).all() and torch.equal(got["input_ids"][0], text)
        return
    got = assemble_inputs(text, am, 999, image_ids, feat, max_vision_token_length=L, check=True)
    want = LO.assemble_inputs(text, am, 999, image_ids, feat, max_vision_token_length=L)
    for k in ("input_ids", "attention_mask", "vision_indices", "coninous_signal"):
        assert torch.equal(got[k], want[k]), k


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 4), st.integers(2, 40), st.integers(0, 2 ** 31 - 1))
def test_get_labels_equals_oracle_on_random_spans(B, T, seed):
    from libra_b200.models.tokenization_libra import get_labels
    from oracle import libra_oracle as LO
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 300, (2, B, T), generator=g)
    ids[:, :, 0] = 1
    boi = 832
    ids[:, :, 1:][torch.rand(2, B, T - 1, generator=g) < 0.1] = boi
    am = torch.ones(B, T, dtype=torch.long)
    rnd = np.random.default_rng(seed)
    spans = []
    for b in range(B):
        if rnd.random() < 0.5:
            am[b, int(rnd.integers(1, T)):] = 0
        sp = []
        for _ in range(int(rnd.integers(0, 3))):
            s = int(rnd.integers(0, T))
            sp.append([s, int(rnd.integers(s, T + 1))])
        spans.append(sp)
    assert torch.equal(get_labels(ids, am, boi, 1, spans), LO.get_labels(ids, am, boi, 1, spans))


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 40), st.booleans()), min_size=1, max_size=12), st.booleans())
def test_adamw_launch_plan_gives_every_element_its_own_decay(params, aligned):
    """FlatAdamW._plan (N3: the reference's decay exclusions, trainer.py:27-37, as ONE launch + a device range table): for every
    element, 'weight decay of the main launch unless inside a no-decay vector range, or redone element-wise as a tail' equals
    the decay of the parameter the element belongs to -- also when run edges are not multiples of the 8-element vector."""
    from libra_b200.optim import FlatAdamW
    runs, lo = [], 0
    for size, decays in params:
        n = size * 8 if aligned else size
        runs.append((lo, lo + n, 0.01 if decays else 0.0))
        lo += n
    p = torch.zeros(lo, dtype=torch.bfloat16)
    opt = FlatAdamW(p, torch.zeros_like(p), weight_decay=0.01, runs=runs)
    wd_main, tbl, tails, n8 = opt._plan()
    eff = torch.full((lo,), float(wd_main))
    if tbl is not None:
        for a, b in tbl.tolist():
            eff[a * 8:b * 8] = 0.0
    redo = torch.zeros(lo, dtype=torch.bool)
    for (sl,) in tails:
        redo[sl] = True
    want = torch.tensor([opt._wd_of(i) for i in range(lo)])
    ok = redo | (eff == want)
    assert ok.all(), (runs, tbl, tails)
    assert (~redo[n8:]).sum() == 0                      # everything behind the last full vector is a tail
    if aligned:
        assert not tails                                 # the model's case: every parameter is a multiple of 8 elements


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 6), st.lists(st.integers(1, 64), min_size=2, max_size=4), st.integers(1, 4096))
def test_flat_gradient_buffer_and_overlapped_pieces_tile_the_buffer(n_layers, widths, min_bytes):
    """FlatGradBuffer lays parameters out in the order their gradients become final (heads, layer L-1 .. 0, embedding side)
    and GradSync's pieces -- issued from the layer hooks, never smaller than min_bytes except at the layer-0 boundary and the
    tail -- tile the buffer in ascending order, each ending on a readiness-group boundary (ONE logical all-reduce)."""
    import torch.distributed as tdist
    from libra_b200 import dist as D

    class Layer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ws = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(w, 3)) for w in widths])

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embed_tokens = torch.nn.Embedding(7, 5)
            self.layers = torch.nn.ModuleList([Layer() for _ in range(n_layers)])
            self.norm = torch.nn.Parameter(torch.zeros(5))
            self.lm_head = torch.nn.Linear(5, 11, bias=False)

    m = Toy()
    buf = D.FlatGradBuffer(m.named_parameters(), fused=False)
    spans = sorted(buf.offsets.values())
    assert spans[0][0] == 0 and spans[-1][1] == buf.numel and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert buf.groups == sorted(buf.groups)
    for n, p in m.named_parameters():
        lo, hi = buf.offsets[n]
        assert p.grad.data_ptr() == buf.flat[lo:hi].data_ptr() and p.grad.shape == p.shape
    # head side first, embedding side last
    assert buf.offsets["lm_head.weight"][1] <= buf.offsets[f"layers.{n_layers - 1}.ws.0"][0]
    assert buf.offsets["embed_tokens.weight"][1] == buf.numel

    class _Done:
        def wait(self):
            return None

    calls = []
    real = tdist.all_reduce
    tdist.all_reduce = lambda t, **kw: (calls.append(t.numel()), _Done())[1]
    try:
        sync = D.GradSync(buf, None, min_bytes=min_bytes, n_layers=n_layers)
        sync.world = 2                                   # pretend: the collective itself is stubbed out
        sync.arm(last=True)
        for li in reversed(range(n_layers)):
            sync.on_layer_grad_ready(li)
        sync.finish()
    finally:
        tdist.all_reduce = real
    pieces = sync.pieces
    assert pieces[0][0] == 0 and pieces[-1][1] == buf.numel and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    assert [hi - lo for lo, hi in pieces] == calls
    ends = set(buf.group_end.values())
    layer0_end = buf.group_end[n_layers]
    for lo, hi in pieces[:-1]:
        assert hi in ends
        assert (hi - lo) * buf.flat.element_size() >= min_bytes or hi == layer0_end
