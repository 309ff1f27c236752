"""Builds libsrk.so (the C-ABI CUDA library, include/srk.h) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libsrk.so")
SOURCES = ["api.cu", "conv_tc.cu", "conv_small.cu", "bandwidth.cu", "wgrad_tc.cu", "metrics.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def _newest_source_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    paths.append(os.path.join(PKG_DIR, "..", "include", "srk.h"))
    return max(os.path.getmtime(p) for p in paths)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libsrk.so unless an up-to-date build exists. Returns the library path."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


def build_trace_library() -> str:
    """Development build with the in-kernel timeline enabled (tools/trace_conv.py); never loaded by the product."""
    out = os.path.join(PKG_DIR, "libsrk_trace.so")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    subprocess.run([nvcc, *NVCC_FLAGS, "-DSRK_TRACE", "-o", out, *[os.path.join(CSRC, s) for s in SOURCES]], check=True)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
