"""Build liblibra_b200.so (sm_100a) in-tree with nvcc.  No fallback targets.

    python -m libra_b200.build [--force]

nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container
(__graft_entry__.build()).  The .so is git-ignored but travels with the tree.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "liblibra_b200.so")
SOURCES = ["host.cu", "norms.cu", "elementwise.cu", "gemm.cu", "gemm_grouped.cu", "gemm_skinny.cu", "attn_fwd.cu", "attn_fwd_stream.cu", "attn_bwd.cu", "attn_bwd_dq_stream.cu", "attn_bwd_dkv.cu", "attn_bwd_dkv_stream.cu", "attn_decode.cu",
           "patch_embed.cu", "vqdec.cu", "preprocess.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _deps(src: str):
    d = [os.path.join(CSRC, src), os.path.join(CSRC, "common.cuh"),
         os.path.join(os.path.dirname(HERE), "include", "libra_b200.h")]
    return [p for p in d if os.path.exists(p)]


def _stale(out: str, deps) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(o, _deps(s)):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[libra_b200.build] compiled {s}")
        return o

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in srcs]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[libra_b200.build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
