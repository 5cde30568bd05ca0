"""ctypes binding of liblibra_b200.so (the C ABI declared in include/libra_b200.h).

The library is the product's only compute path: if it is missing or the device is not
sm_100 every op raises -- there is no eager/CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

import torch  # noqa: F401  (loads libcudart.so.12 that the library links against)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblibra_b200.so")

P, I, L, F = c_void_p, c_int, c_int64, c_float

# name -> (restype, argtypes); mirrors include/libra_b200.h one to one
SIGNATURES = {
    "lb_version": (I, []),
    "lb_last_error": (I, [c_char_p, I]),
    "lb_device_check": (I, []),
    "lb_sm_count": (I, []),
    "lb_set_pdl": (I, [I]),
    "lb_gemm_skinny_set_trace": (I, [P]),
    "lb_clip_preprocess_workspace": (L, [P, P, I, I, I, I]),
    "lb_clip_resample_coeffs": (I, [I, I, I, I, P, I, P]),
    "lb_clip_preprocess": (I, [P, P, P, P, I, I, I, I, P, P, P, ctypes.c_double, P, I, P, P, L, P]),
    "lb_rmsnorm_fwd": (I, [P, P, P, P, P, P, L, I, F, P]),
    "lb_rmsnorm_bwd_workspace": (L, [L, I]),
    "lb_rmsnorm_bwd": (I, [P, P, P, P, P, P, P, P, P, P, P, L, I, P]),
    "lb_layernorm_fwd": (I, [P, P, P, P, P, P, L, I, F, P]),
    "lb_layernorm_bwd_workspace": (L, [L, I]),
    "lb_layernorm_bwd": (I, [P, P, P, P, P, P, P, P, P, L, I, P]),
    "lb_swiglu_fwd": (I, [P, P, P, L, I, L, L, L, P]),
    "lb_swiglu_bwd": (I, [P, P, P, P, P, L, I, L, L, L, L, L, P]),
    "lb_bias_quick_gelu_fwd": (I, [P, P, P, L, I, P]),
    "lb_bias_quick_gelu_bwd": (I, [P, P, P, P, L, I, P]),
    "lb_gather_rows": (I, [P, P, P, L, I, P]),
    "lb_embed_lang_fwd": (I, [P, P, P, L, I, P]),
    "lb_embed_vision_cat_fwd": (I, [P, P, P, P, P, P, P, L, I, I, P]),
    "lb_embed_bwd": (I, [P, P, L, I, P, L, I, P]),
    "lb_lfq_pack": (I, [P, I, L, I, I, I, L, L, L, P, P]),
    "lb_lfq_unpack": (I, [P, L, I, I, P, I, P]),
    "lb_attn_prep_fwd": (I, [P] * 15 + [L, I, I, P, P]),
    "lb_attn_prep_fwd_bridge": (I, [P] * 9 + [I] + [P] * 10 + [L, I, I, P, P]),
    "lb_attn_prep_bwd": (I, [P] * 15 + [L, I, I, P]),
    "lb_attn_fwd": (I, [P, P, P, P, P, P, P, I, P, P, P, P, P, I, I, I, I, I, F, P]),
    "lb_attn_fwd_stream": (I, [P, P, P, P, P, P, P, I, P, P, I, I, I, P, P, P, P, P, I, I, I, I, I, F, P]),
    "lb_attn_fwd_stream_max_cta_items": (I, []),
    "lb_attn_fwd_stream_set_cta_log": (I, [P]),
    "lb_attn_bwd_dkv_set_cta_log": (I, [P]),
    "lb_attn_decode_workspace_floats": (I, [I, I, I, I]),
    "lb_attn_decode": (I, [P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, F, P]),
    "lb_attn_fwd_stream_set_trace": (I, [P]),
    "lb_attn_fwd_set_trace": (I, [P]),
    "lb_attn_bwd_prepare": (I, [P, P, P, P, P, I, I, I, I, P]),
    "lb_attn_bwd_dq": (I, [P] * 10 + [I, P, P, P, I, I, I, I, I, F, P]),
    "lb_attn_bwd_dq_stream": (I, [P] * 10 + [I, P, P, I, I, I, P, P, P, I, I, I, I, I, F, P]),
    "lb_attn_bwd_dq_stream_max_cta_items": (I, []),
    "lb_attn_bwd_dq_stream_set_trace": (I, [P]),
    "lb_attn_bwd_dkv_stream": (I, [P] * 11 + [I, P, P, I, I, I, P, P, P, P, P, P, I, I, I, I, I, F, P]),
    "lb_attn_bwd_dkv_stream_max_cta_items": (I, []),
    "lb_attn_bwd_dkv_stream_supported": (I, []),
    "lb_attn_bwd_dkv": (I, [P] * 11 + [I, P, P, P, P, P, P, I, I, I, I, I, F, P]),
    "lb_gemm_bf16": (I, [P, P, P, P, L, L, L, L, L, L, I, I, I, I, I, P]),
    "lb_gemm_grouped_workspace_bytes": (I, [P, I]),
    "lb_gemm_grouped": (I, [P, I, P, L, P]),
    "lb_gemm_skinny_workspace_bytes": (L, [P, I]),
    "lb_gemm_skinny": (I, [P, I, P, L, P]),
    "lb_gemm_tmap_cache_stats": (I, [P, P]),
    "lb_patch_embed_pack_weight": (I, [P, P, I, I, P]),
    "lb_patch_embed_fwd": (I, [P, P, P, P, P, I, I, I, I, P]),
    "lb_cross_entropy_fwd_bwd": (I, [P, L, P, P, L, I, F, P]),
    "lb_probe_umma": (I, [I, P, P, P, I, P]),
    "lb_adamw_bf16": (I, [P, P, P, P, L, F, F, F, F, F, I, P]),
    "lb_adamw_bf16_scaled": (I, [P, P, P, P, L, F, F, F, F, F, I, P, P, I, P]),
    "lb_grad_clip_scale": (I, [P, L, F, P, I, P, P]),
    "lb_vq_codes": (I, [P, L, I, L, I, P, I, P]),
    "lb_vq_groupnorm_chunks": (I, [I, I]),
    "lb_vq_groupnorm": (I, [P, P, P, P, P, I, I, I, I, I, F, I, I, I, P]),
    "lb_vq_upsample_nearest": (I, [P, P, P, P, I, I, I, I, I, I, I, P]),
    "lb_vq_pad": (I, [P, L, P, P, I, I, I, I, P]),
    "lb_vq_to_nchw": (I, [P, P, I, I, I, I, I, P]),
    "lb_softmax_rows": (I, [P, L, I, L, F, P]),
}

_lib = None


class LibraB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (building is the job of libra_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraB200Error(
            f"{LIB_PATH} not found: build it with `python -m libra_b200.build` "
            "(libra_b200 has no fallback path without its CUDA library)")
    lib = ctypes.CDLL(LIB_PATH)
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    lib._missing = missing
    _lib = lib
    return lib


def last_error() -> str:
    lib = load()
    buf = ctypes.create_string_buffer(512)
    lib.lb_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


# CUDA kernels enqueued by one successful call of each entry point (host-side launch accounting for bench.py)
KERNELS_PER_CALL = {
    "lb_rmsnorm_fwd": 1, "lb_rmsnorm_bwd": 2, "lb_layernorm_fwd": 1, "lb_layernorm_bwd": 2, "lb_swiglu_fwd": 1,
    "lb_swiglu_bwd": 1, "lb_bias_quick_gelu_fwd": 1, "lb_bias_quick_gelu_bwd": 1, "lb_gather_rows": 1,
    "lb_embed_lang_fwd": 1, "lb_embed_vision_cat_fwd": 3, "lb_embed_bwd": 1, "lb_lfq_pack": 1, "lb_lfq_unpack": 1,
    "lb_attn_prep_fwd": 1, "lb_attn_prep_fwd_bridge": 1, "lb_attn_prep_bwd": 1, "lb_attn_fwd": 1, "lb_attn_fwd_stream": 1, "lb_attn_bwd_prepare": 1, "lb_attn_bwd_dq": 1, "lb_attn_bwd_dq_stream": 1,
    "lb_attn_bwd_dkv": 1, "lb_attn_bwd_dkv_stream": 1, "lb_attn_decode": 2, "lb_gemm_bf16": 1, "lb_gemm_grouped": 1, "lb_gemm_skinny": 1, "lb_patch_embed_fwd": 1, "lb_patch_embed_pack_weight": 1, "lb_cross_entropy_fwd_bwd": 1, "lb_probe_umma": 1, "lb_adamw_bf16": 1, "lb_adamw_bf16_scaled": 1, "lb_grad_clip_scale": 2,
    "lb_clip_preprocess": 2,
}
launch_counts: dict = {}


def reset_launch_counts():
    launch_counts.clear()


def total_launches() -> int:
    return sum(KERNELS_PER_CALL.get(k, 1) * v for k, v in launch_counts.items())


def call(name: str, *args):
    """Invoke an int-returning entry point; raise with the library's message on failure."""
    lib = load()
    fn = getattr(lib, name)
    rc = fn(*args)
    launch_counts[name] = launch_counts.get(name, 0) + 1
    if rc != 0:
        raise LibraB200Error(f"{name} failed ({rc}): {last_error()}")
    return rc


def require_device():
    """Fail loudly unless a CUDA sm_100 device is current."""
    if not torch.cuda.is_available():
        raise LibraB200Error("libra_b200 needs a CUDA sm_100 (B200) device; none is visible and there is no CPU path")
    call("lb_device_check")
