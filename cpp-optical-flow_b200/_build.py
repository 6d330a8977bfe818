"""Build recipe for libhs_b200.so (hand-written sm_100a kernels + the C ABI of include/hs.h).

The library is built IN-TREE next to this file so that it travels with the repo snapshot to the
GPU box; it is git-ignored (*.so).  nvcc cross-compiles for sm_100a without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhs_b200.so")
SOURCES = [os.path.join(CSRC, "hs_api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "hs_kernels.cuh"), os.path.join(CSRC, "hs_multi.inl"), os.path.join(CSRC, "hs_kernels_f64.cuh"),
                  os.path.join(os.path.dirname(HERE), "include", "hs.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def nvcc_path() -> str | None:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources; return its path."""
    if not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: libhs_b200.so cannot be built (and there is no CPU fallback)")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    env = dict(os.environ)
    # the image's CC/CXX wrappers are not needed by nvcc; use the system host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB
