"""Builds libpixlzr_b200.so (sm_100a only) in-tree with nvcc.

    python pixlzr-rust_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
snapshot.  cudart is linked statically; NCCL is dlopen'ed lazily at run time.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("PXZ_OUT") or os.path.join(HERE, "libpixlzr_b200.so")
SOURCES = ["kernels.cu", "qoi_device.cu", "abi.cpp", "tables.cpp", "container.cpp", "nccl_dyn.cpp"]
HEADERS = ["pxz_internal.h", "pxz_host.h", "srgb_lut.inc", "resample_warp.cuh", "resample_tma.cuh", "analyze_sobel_tma.cuh", os.path.join("..", "..", "include", "pixlzr_b200.h")]
NVCC = os.environ.get("PXZ_NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", "/usr/bin/g++",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-shared", "-cudart", "static",
]


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    cmd = [NVCC] + FLAGS + os.environ.get("PXZ_EXTRA_NVCC_FLAGS", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libpixlzr_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
