"""Correctness + timing of the grouped tcgen05 GEMM (csrc/gemm_grouped.cu) against torch (cuBLAS) on the decoder's shapes.

    LB_GEMM_CG=2 python scripts/gemm_bench.py [--quick] [--out gpurun_out/gemm_cg2.json]

Every case prints one JSON line; a summary table goes to stdout at the end.  The cuBLAS column is the bar the
round-1 verdict set ("beat nvjet on the same box"); it is printed beside ours, never substituted for it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libra_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
dev = "cuda"


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, dtype=torch.float32) * scale).to(BF16)


def ref_mm(a, b, ta, tb):
    A = a.float().t() if ta else a.float()
    Bm = b.float() if tb else b.float().t()
    return A @ Bm


def check(name, got, want, tol=2e-2):
    got, want = got.float(), want.float()
    denom = want.abs().max().item() + 1e-6
    err = (got - want).abs().max().item() / denom
    bad = not (err < tol) or not torch.isfinite(got).all().item()
    print(json.dumps({"case": name, "rel_err": round(err, 6), "ok": not bad}), flush=True)
    return not bad


def time_fn(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3      # us


def correctness(quick):
    ok = True
    torch.manual_seed(0)
    cases = [
        # M, N, K, ta, tb
        (256, 256, 64, 0, 0), (256, 256, 256, 0, 0), (300, 200, 136, 0, 0), (1000, 520, 1000, 0, 0),
        (128, 64, 64, 0, 0), (5880, 4096, 512, 0, 0), (2312, 1024, 4096, 0, 0),
        (512, 384, 320, 0, 1), (512, 384, 320, 1, 0), (512, 384, 320, 1, 1),
        (1000, 776, 520, 0, 1), (1000, 776, 520, 1, 0), (776, 1000, 264, 1, 1),
        (600, 8, 512, 0, 0), (600, 512, 8, 0, 0), (600, 8, 512, 0, 1), (520, 16, 600, 1, 0), (8, 512, 600, 1, 1),
        (700, 514, 1024, 0, 0), (514, 1024, 700, 1, 1), (700, 1024, 514, 0, 1), (700, 1024, 514, 0, 0), (300, 514, 200, 0, 1),
    ]
    if quick:
        cases = cases[:8]
    def padded(rows, cols, scale=1.0):
        """[rows, cols] view with a row pitch rounded up to 8 elements, NaN in the padding columns"""
        ld = (cols + 7) // 8 * 8
        buf = torch.full((rows, ld), float("nan"), device=dev, dtype=BF16)
        buf[:, :cols] = rnd(rows, cols, scale=scale)
        return buf[:, :cols]

    for (M, N, K, ta, tb) in cases:
        a = padded(K, M) if ta else padded(M, K)
        b = padded(K, N) if tb else padded(N, K)
        ldc = (N + 7) // 8 * 8
        cbuf = torch.full((M + 3, ldc), float("nan"), device=dev, dtype=BF16)      # 3 guard rows behind the tensor
        c = cbuf[:M, :N]
        ops.gemm_grouped([ops.gp(a, b, c, ta=bool(ta), tb=bool(tb))])
        ok &= check(f"plain M{M} N{N} K{K} ta{ta} tb{tb}", c, ref_mm(a, b, ta, tb) / 1.0)
        if not torch.isnan(cbuf[M:].float()).all():
            print(json.dumps({"case": f"guard rows M{M} N{N}", "ok": False}), flush=True)
            ok = False
        if ldc > N:
            pad = cbuf[:M, N:].float()
            clean = bool((torch.isnan(pad) | (pad == 0)).all())
            print(json.dumps({"case": f"padding columns M{M} N{N}", "untouched": bool(torch.isnan(pad).all()), "zero_or_untouched": clean}), flush=True)
            ok &= clean
    # addend (out of place and in place), bias + quick_gelu with pre-activation, SwiGLU
    M, N, K = 1000, 768, 512
    a, b, d = rnd(M, K), rnd(N, K, scale=0.05), rnd(M, N)
    c = torch.empty(M, N, device=dev, dtype=BF16)
    ops.gemm_grouped([ops.gp(a, b, c, d=d)])
    want = (a.float() @ b.float().t()).to(BF16).float() + d.float()
    ok &= check("addend out-of-place", c, want)
    c2 = d.clone()
    ops.gemm_grouped([ops.gp(a, b, c2, d=c2)])
    ok &= check("addend in-place", c2, want)
    bias = rnd(N)
    pre = torch.empty(M, N, device=dev, dtype=BF16)
    ops.gemm_grouped([ops.gp(a, b, c, bias=bias, epi=ops.EPI_QGELU, g=pre)])
    t = (a.float() @ b.float().t() + bias.float()).to(BF16).float()
    ok &= check("bias pre-activation", pre, t)
    ok &= check("bias quick_gelu", c, t * torch.sigmoid(1.702 * t))
    ops.gemm_grouped([ops.gp(a, b, c, bias=bias)])
    ok &= check("bias only", c, t)
    for (M, N, K) in [(600, 384, 256), (1000, 2752, 512), (300, 11008 // 8, 320)]:
        a, wg, wu = rnd(M, K), rnd(N, K, scale=0.05), rnd(N, K, scale=0.05)
        h, g, u = (torch.empty(M, N, device=dev, dtype=BF16) for _ in range(3))
        ops.gemm_grouped([ops.gp(a, wg, h, b2=wu, epi=ops.EPI_SWIGLU, g=g, u=u)])
        gr = (a.float() @ wg.float().t()).to(BF16)
        ur = (a.float() @ wu.float().t()).to(BF16)
        hr = torch.nn.functional.silu(gr).float() * ur.float()
        ok &= check(f"swiglu gate M{M} N{N}", g, gr)
        ok &= check(f"swiglu up M{M} N{N}", u, ur)
        ok &= check(f"swiglu out M{M} N{N}", h, hr)
        h2 = torch.empty_like(h)
        ops.gemm_grouped([ops.gp(a, wg, h2, b2=wu, epi=ops.EPI_SWIGLU)])
        ok &= check(f"swiglu out only M{M} N{N}", h2, hr)
    # grouped: routed projection = dense language rows | chained low-rank vision rows, one launch
    nl, nv, H, R = 1000, 600, 1024, 256
    x = rnd(nl + nv, H)
    W, A, Bw = rnd(H, H, scale=0.03), rnd(R, H, scale=0.03), rnd(H, R, scale=0.03)
    y = torch.empty(nl + nv, H, device=dev, dtype=BF16)
    mid = torch.empty(nv, R, device=dev, dtype=BF16)
    res = rnd(nl + nv, H)
    for rep in range(3):
        y.fill_(float("nan")); mid.fill_(float("nan"))
        ops.gemm_grouped([ops.gp(x[nl:], A, mid), ops.gp(x[:nl], W, y[:nl], d=res[:nl]),
                          ops.gp(mid, Bw, y[nl:], d=res[nl:], wait_on=0)])
        midr = (x[nl:].float() @ A.float().t()).to(BF16)
        yr = torch.cat([(x[:nl].float() @ W.float().t()).to(BF16).float() + res[:nl].float(),
                        (midr.float() @ Bw.float().t()).to(BF16).float() + res[nl:].float()])
        ok &= check(f"grouped chain rep{rep} mid", mid, midr)
        ok &= check(f"grouped chain rep{rep} y", y, yr)
    # K segments summed in the accumulator (fan-out dgrad), transposed dependent (wgrad of a chain), alpha
    M, H2 = 1100, 768
    dq, dk, dt = rnd(M, 512), rnd(M, 384), rnd(M, 8)
    Wq, Wk, Al = rnd(512, H2, scale=0.05), rnd(384, H2, scale=0.05), rnd(8, H2, scale=0.05)
    dx = torch.full((M, H2), float("nan"), device=dev, dtype=BF16)
    ops.gemm_grouped([ops.gp(dq, Wq, dx, tb=True), ops.gp(dk, Wk, dx, tb=True, acc_prev=True), ops.gp(dt, Al, dx, tb=True, acc_prev=True)])
    ok &= check("3 K-segments", dx, dq.float() @ Wq.float() + dk.float() @ Wk.float() + dt.float() @ Al.float())
    dy, Bw2, xv = rnd(M, H2), rnd(H2, 256, scale=0.05), rnd(M, 640)
    dmid = torch.full((M, 256), float("nan"), device=dev, dtype=BF16)
    dA = torch.full((256, 640), float("nan"), device=dev, dtype=BF16)
    dxv = torch.full((M, 640), float("nan"), device=dev, dtype=BF16)
    A2 = rnd(256, 640, scale=0.05)
    alpha = torch.tensor([0.5], device=dev, dtype=torch.float32)
    for rep in range(2):
        ops.gemm_grouped([ops.gp(dy, Bw2, dmid, tb=True),
                          ops.gp(dmid, xv, dA, ta=True, tb=True, wait_on=0, alpha=alpha),
                          ops.gp(dmid, A2, dxv, tb=True, wait_on=0)])
        dmr = (dy.float() @ Bw2.float()).to(BF16).float()
        ok &= check(f"transposed dependent (whole-problem wait) rep{rep}", dA, 0.5 * (dmr.t() @ xv.float()))
        ok &= check(f"row-block dependent rep{rep}", dxv, dmr @ A2.float())
    # empty segments are skipped
    ops.gemm_grouped([ops.gp(x[:0], W, y[:0]), ops.gp(x[:64], W, y[:64])])
    ok &= check("empty segment", y[:64], x[:64].float() @ W.float().t())
    # many tiles per CTA + both accumulator buffers + ring wrap
    M, N, K = 4096, 4096, 1088
    a, b = rnd(M, K), rnd(N, K, scale=0.05)
    c = torch.empty(M, N, device=dev, dtype=BF16)
    ops.gemm_grouped([ops.gp(a, b, c)])
    ok &= check("multi-wave 4096x4096x1088", c, a.float() @ b.float().t())
    return ok


def bench(quick):
    rows = []
    nl, nv, H, I = 5880, 2312, 4096, 11008
    shapes = [
        ("lang fwd qkv/o  NT", nl, H, H, 0, 0),
        ("lang fwd gate/up NT", nl, I, H, 0, 0),
        ("lang fwd down NT", nl, H, I, 0, 0),
        ("lang dgrad NN (dy W)", nl, H, H, 0, 1),
        ("lang dgrad down NN", nl, I, H, 0, 1),
        ("lang wgrad TN 4096x4096", H, H, nl, 1, 1),
        ("lang wgrad TN 11008x4096", I, H, nl, 1, 1),
        ("vis stage1 NT", nv, 1024, H, 0, 0),
        ("vis stage2 NT", nv, H, 1024, 0, 0),
        ("lm_head NT", nl, 32000, H, 0, 0),
        ("square 8192", 8192, 8192, 8192, 0, 0),
    ]
    if quick:
        shapes = shapes[:3] + shapes[5:6] + shapes[-1:]
    for name, M, N, K, ta, tb in shapes:
        a = rnd(K, M) if ta else rnd(M, K)
        b = rnd(K, N, scale=0.05) if tb else rnd(N, K, scale=0.05)
        c = torch.empty(M, N, device=dev, dtype=BF16)
        A_ = a.t() if ta else a
        B_ = b if tb else b.t()
        t_ours = time_fn(lambda: ops.gemm_grouped([ops.gp(a, b, c, ta=bool(ta), tb=bool(tb))]))
        c2 = torch.empty(M, N, device=dev, dtype=BF16)
        t_cublas = time_fn(lambda: torch.matmul(A_, B_, out=c2))
        fl = 2.0 * M * N * K
        err = ((c.float() - c2.float()).abs().max() / (c2.float().abs().max() + 1e-6)).item()
        row = {"bench": name, "M": M, "N": N, "K": K, "ta": ta, "tb": tb, "us_ours": round(t_ours, 1),
               "us_cublas": round(t_cublas, 1), "tf_ours": round(fl / t_ours / 1e6, 1), "tf_cublas": round(fl / t_cublas / 1e6, 1),
               "ratio": round(t_cublas / t_ours, 3), "rel_diff_vs_cublas": round(err, 5)}
        print(json.dumps(row), flush=True)
        rows.append(row)
    # the grouped launches of one decoder layer (forward): q/k/v fan-out, o-proj with residual, gate|up SwiGLU, down
    x = rnd(nl + nv, H)
    Wq, Wk, Wv = (rnd(H, H, scale=0.02) for _ in range(3))
    Aq, Ak, Av = (rnd(1024, H, scale=0.02) for _ in range(3))
    Bq, Bk, Bv = (rnd(H, 1024, scale=0.02) for _ in range(3))
    q, k, v = (torch.empty(nl + nv, H, device=dev, dtype=BF16) for _ in range(3))
    mq, mk, mv = (torch.empty(nv, 1024, device=dev, dtype=BF16) for _ in range(3))

    def fanout_ours():
        ops.gemm_grouped([ops.gp(x[nl:], Aq, mq), ops.gp(x[nl:], Ak, mk), ops.gp(x[nl:], Av, mv),
                          ops.gp(x[:nl], Wq, q[:nl]), ops.gp(x[:nl], Wk, k[:nl]), ops.gp(x[:nl], Wv, v[:nl]),
                          ops.gp(mq, Bq, q[nl:], wait_on=0), ops.gp(mk, Bk, k[nl:], wait_on=1), ops.gp(mv, Bv, v[nl:], wait_on=2)])

    side = torch.cuda.Stream()

    def fanout_cublas():
        main = torch.cuda.current_stream()
        side.wait_event(main.record_event())
        with torch.cuda.stream(side):
            for A_, B_, m_, y_ in ((Aq, Bq, mq, q), (Ak, Bk, mk, k), (Av, Bv, mv, v)):
                torch.matmul(x[nl:], A_.t(), out=m_)
                torch.matmul(m_, B_.t(), out=y_[nl:])
        for W_, y_ in ((Wq, q), (Wk, k), (Wv, v)):
            torch.matmul(x[:nl], W_.t(), out=y_[:nl])
        main.wait_event(side.record_event())

    fl = 3 * (2.0 * nl * H * H + 2.0 * nv * H * 1024 * 2)
    t1, t2 = time_fn(fanout_ours), time_fn(fanout_cublas)
    row = {"bench": "layer fwd q/k/v fan-out (9 problems, 1 launch) vs cuBLAS 2 streams", "us_ours": round(t1, 1), "us_cublas": round(t2, 1),
           "tf_ours": round(fl / t1 / 1e6, 1), "tf_cublas": round(fl / t2 / 1e6, 1), "ratio": round(t2 / t1, 3)}
    print(json.dumps(row), flush=True)
    rows.append(row)
    # gate|up with the SwiGLU epilogue vs cuBLAS gate, up + the separate swiglu pass
    Wg, Wu = rnd(I, H, scale=0.02), rnd(I, H, scale=0.02)
    Ag, Au = rnd(2752, H, scale=0.02), rnd(2752, H, scale=0.02)
    Bg, Bu = rnd(I, 2752, scale=0.02), rnd(I, 2752, scale=0.02)
    g, u, h = (torch.empty(nl + nv, I, device=dev, dtype=BF16) for _ in range(3))
    mg, mu = (torch.empty(nv, 2752, device=dev, dtype=BF16) for _ in range(2))

    def mlp_up_ours():
        ops.gemm_grouped([ops.gp(x[nl:], Ag, mg), ops.gp(x[nl:], Au, mu),
                          ops.gp(x[:nl], Wg, h[:nl], b2=Wu, epi=ops.EPI_SWIGLU, g=g[:nl], u=u[:nl]),
                          ops.gp(mg, Bg, g[nl:], wait_on=0), ops.gp(mu, Bu, u[nl:], wait_on=1)])
        ops.swiglu_fwd(g[nl:], u[nl:], out=h[nl:])

    def mlp_up_cublas():
        main = torch.cuda.current_stream()
        side.wait_event(main.record_event())
        with torch.cuda.stream(side):
            for A_, B_, m_, y_ in ((Ag, Bg, mg, g), (Au, Bu, mu, u)):
                torch.matmul(x[nl:], A_.t(), out=m_)
                torch.matmul(m_, B_.t(), out=y_[nl:])
        for W_, y_ in ((Wg, g), (Wu, u)):
            torch.matmul(x[:nl], W_.t(), out=y_[:nl])
        main.wait_event(side.record_event())
        ops.swiglu_fwd(g, u)

    fl = 2 * (2.0 * nl * H * I + 2.0 * nv * H * 2752 + 2.0 * nv * 2752 * I)
    t1, t2 = time_fn(mlp_up_ours), time_fn(mlp_up_cublas)
    row = {"bench": "layer fwd gate|up (+SwiGLU epilogue on language rows) vs cuBLAS + swiglu pass", "us_ours": round(t1, 1),
           "us_cublas": round(t2, 1), "tf_ours": round(fl / t1 / 1e6, 1), "tf_cublas": round(fl / t2 / 1e6, 1), "ratio": round(t2 / t1, 3)}
    print(json.dumps(row), flush=True)
    rows.append(row)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-bench", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    t0 = time.time()
    ok = correctness(args.quick)
    print(json.dumps({"correctness_all_ok": bool(ok), "cg": os.environ.get("LB_GEMM_CG", "2"), "s": round(time.time() - t0, 1)}), flush=True)
    rows = [] if args.no_bench else bench(args.quick)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"ok": bool(ok), "cg": os.environ.get("LB_GEMM_CG", "2"), "rows": rows,
                       "gpu": torch.cuda.get_device_name(0)}, f, indent=1)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
