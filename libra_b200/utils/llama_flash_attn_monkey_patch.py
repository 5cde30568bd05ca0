"""Counterpart of the reference's utils/llama_flash_attn_monkey_patch.py (:75-178 the FlashAttention-2 forward, :190-202 the
hook).  In the reference the hook is disabled (`replace_llama_attn_with_flash_attn()` raises NotImplementedError, :191) because
FA2 cannot express Libra's bridge: S_ij = q_i.(k_j + [flag_i != flag_j] kb_j), O_i = sum_j P_ij (v_j + [flag_i != flag_j] vb_j).

Here LibraAttention already runs on this library's fused flash-attention kernels WITH the bridge (csrc/attn_fwd_stream.cu,
attn_bwd*.cu), so the hook has nothing to replace: it verifies that the CUDA path is usable and returns.  A training script
that calls it (utils/train_utils.py:22-23, commented out in the reference) keeps working unmodified.

`flash_attn_reference_point` is the informational FA2 timing point SURVEY.md section 8 (A11) asks for: the flash-attn library
on the same shape with use_bridge=False semantics (plain causal attention: numerically NOT Libra's attention)."""
from __future__ import annotations

import warnings


def replace_llama_attn_with_flash_attn():
    import torch
    from .. import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("libra_b200's fused bridge attention needs a CUDA sm_100 device")
    _lib.require_device()
    major, _ = torch.cuda.get_device_capability()
    if major != 10:
        warnings.warn("libra_b200 attention kernels are built for sm_100a only")
    return True


def flash_attn_reference_point(batch=4, seqlen=2048, heads=32, head_dim=128, iters=10):
    """(forward ms, forward+backward ms) of flash_attn_func(causal=True) on [B,T,H,D] bf16, or None if the library cannot run
    on this device.  Informational only: no bridge, so it is not a parity reference."""
    import torch
    try:
        from flash_attn import flash_attn_func
    except Exception:
        return None
    try:
        g = torch.Generator(device="cuda").manual_seed(0)
        q, k, v = (torch.randn(batch, seqlen, heads, head_dim, device="cuda", generator=g).bfloat16().requires_grad_(True) for _ in range(3))

        def t(fn):
            for _ in range(3):
                fn()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                fn()
            e.record()
            torch.cuda.synchronize()
            return s.elapsed_time(e) / iters

        f = t(lambda: flash_attn_func(q, k, v, causal=True))
        fb = t(lambda: flash_attn_func(q, k, v, causal=True).float().sum().backward())
        return f, fb
    except Exception as ex:          # library built without kernels for this architecture
        return ("unavailable", repr(ex)[:200])
