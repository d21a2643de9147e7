"""Build recipe of libgibbs_b200.so (sm_100a only, in-tree so the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libgibbs_b200.so")
SOURCES = [os.path.join(CSRC, "gibbs_api.cu"), os.path.join(CSRC, "gibbs_drift_launch.cu")]
DEPS = SOURCES + [
    os.path.join(CSRC, "gibbs_device.cuh"),
    os.path.join(CSRC, "gibbs_kernels.cuh"),
    os.path.join(CSRC, "gibbs_motif.cuh"),
    os.path.join(CSRC, "gibbs_drift.cuh"),
    os.path.join(CSRC, "gibbs_drift_dev.cuh"),
    os.path.join(ROOT, "include", "gibbs_b200.h"),
]

NVCC_FLAGS = [
    "--threads", "0",         # the translation units in parallel
    "--split-compile", "0",   # parallel ptxas over the template instantiations
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # float64 products must round like the reference: no FMA contraction
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library with nvcc (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
