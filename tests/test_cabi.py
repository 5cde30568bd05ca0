"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/libra_b200.h declares,
with the argument counts the ctypes binding assumes.  No compute calls (runs without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _prototypes():
    src = open(os.path.join(ROOT, "include", "libra_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t)\s+(lb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(1)] = n
    return protos


@pytest.fixture(scope="module")
def lib():
    from libra_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_declares_entry_points():
    protos = _prototypes()
    assert len(protos) >= 25
    for must in ("lb_attn_fwd", "lb_attn_bwd_dq", "lb_attn_bwd_dkv", "lb_rmsnorm_fwd", "lb_lfq_pack", "lb_gemm_bf16",
                 "lb_patch_embed_fwd", "lb_cross_entropy_fwd_bwd"):
        assert must in protos


def test_library_exports_every_declared_symbol(lib):
    protos = _prototypes()
    missing = [n for n in protos if not hasattr(lib, n)]
    assert not missing, f"declared in include/libra_b200.h but not exported: {missing}"


def test_ctypes_signatures_match_header(lib):
    from libra_b200 import _lib
    protos = _prototypes()
    assert set(_lib.SIGNATURES) == set(protos)
    for name, (_, args) in _lib.SIGNATURES.items():
        assert len(args) == protos[name], (name, len(args), protos[name])


def test_version_and_error_text(lib):
    import ctypes
    assert lib.lb_version() >= 100
    buf = ctypes.create_string_buffer(64)
    assert lib.lb_last_error(buf, 64) == 0


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    from libra_b200 import _lib
    # argument validation happens before any CUDA call
    rc = lib.lb_rmsnorm_fwd(None, None, None, None, None, None, 4, 12, 1e-6, None)   # cols % 8 != 0
    assert rc == -1 and "cols" in _lib.last_error()
    rc = lib.lb_attn_fwd(None, None, None, None, None, None, None, 1, None, None, None, None, None, 1, 128, 2, 96, 1, 1.0, None)
    assert rc == -1 and "head_dim" in _lib.last_error()


def test_product_fails_loudly_without_cuda():
    import torch
    from libra_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.LibraB200Error):
        _lib.require_device()
