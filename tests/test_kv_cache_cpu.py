"""Host logic of the decode path's KV cache (N1) on CPU tensors: prefill + one-token appends through the device-side length,
capacity growth, and the conversion to the reference's per-layer tuple (modeling_libra.py:354-361)."""
import torch

from libra_b200.kv_cache import LibraKVCache


def test_append_commit_reserve_and_reference_layout():
    torch.manual_seed(0)
    L, B, H, D, T0 = 2, 3, 2, 8, 5
    C = H * D
    cache = LibraKVCache(L, B, T0 + 2, H, D, "cpu")
    mk = lambda n: torch.randn(B * n, C).bfloat16()
    hist = {k: [[] for _ in range(L)] for k in ("k_fv", "k_fl", "v_fv", "v_fl")}
    flags = []

    def step(q_len, flag):
        for layer in range(L):
            t = {k: mk(q_len) for k in hist}
            cache.append(layer, t["k_fv"], t["k_fl"], t["v_fv"], t["v_fl"], q_len)
            for k in hist:
                hist[k][layer].append(t[k].view(B, q_len, C))
        cache.commit(flag)
        flags.append(flag)

    step(T0, torch.rand(B, T0) > 0.5)                       # prefill
    assert cache.get_seq_length() == T0 and int(cache.len_dev) == T0
    for _ in range(6):                                       # one-token steps, past the initial capacity
        cache.reserve(1)
        step(1, torch.rand(B, 1) > 0.5)
    T = T0 + 6
    assert cache.get_seq_length() == T and int(cache.len_dev) == T and cache.capacity >= T
    flag = torch.cat(flags, dim=1)
    assert torch.equal(cache.flag[:, :T], flag)
    for layer in range(L):
        for k in hist:
            want = torch.cat(hist[k][layer], dim=1)
            assert torch.equal(getattr(cache, k)[layer][:, :T], want), (k, layer)
    ref = cache.to_reference()
    assert len(ref) == L == len(cache)
    (kfv, kfl), v, vb, f = ref[1]
    assert kfv.shape == (B, H, T, D) and v.shape == vb.shape == (B, H, T, D) and torch.equal(f, flag)
    hd = lambda t: t.view(B, T, H, D).transpose(1, 2)
    v_fv, v_fl = hd(torch.cat(hist["v_fv"][1], dim=1)), hd(torch.cat(hist["v_fl"][1], dim=1))
    fk = flag[:, None, :, None]
    # a key's own-modality value is the plain one; the other variant carries the bridge
    assert torch.equal(v, torch.where(fk, v_fv, v_fl))
    assert torch.allclose((v.float() + vb.float()), torch.where(fk, v_fl, v_fv).float(), atol=2e-2, rtol=2e-2)
    assert torch.equal(kfl, hd(torch.cat(hist["k_fl"][1], dim=1)))
