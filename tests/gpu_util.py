import pytest
import torch


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from libra_b200 import _lib
    _lib.require_device()      # fails loudly (never skips) if the CUDA library is missing on a GPU box


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def assert_close(got, want, rtol=2e-2, atol=2e-2, msg=""):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    lim = atol + rtol * want.abs()
    bad = (err > lim)
    assert not bad.any(), f"{msg} max err {err.max().item():.4g} at {bad.nonzero()[:3].tolist()} ({bad.sum().item()} bad of {bad.numel()}), rel_fro={rel_err(got, want):.3g}"
