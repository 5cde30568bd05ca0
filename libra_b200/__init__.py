"""libra_b200 -- B200-native (sm_100a) kernels and host mirror for the Libra training hot path.

Layout: csrc/ (CUDA kernels + C ABI, built into liblibra_b200.so), _lib.py (ctypes binding),
ops.py (raw ops), schedule.py (routing permutation and attention work lists), functional.py
(autograd functions), models/ (host-side mirror of the reference's module interface).
"""
__version__ = "0.1.0"
