"""Builds libsrk.so (the C-ABI CUDA library, include/srk.h) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libsrk.so")
SOURCES = ["api.cu", "conv_tc.cu", "conv_small.cu", "bandwidth.cu", "wgrad_tc.cu", "metrics.cu", "espcn_fused.cu", "espcn_fused_c1.cu", "espcn_fused_c3.cu", "collective.cu", "gemm_tc.cu", "f2_ops.cu", "conv_strip.cu", "peer_reduce.cu"]
NVCC_FLAGS_COMPILE = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"] + os.environ.get("SRK_NVCC_EXTRA", "").split()
NVCC_FLAGS_LINK = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]


def _header_paths() -> list[str]:
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(PKG_DIR, "..", "include", "srk.h")]


def _newest_source_mtime() -> float:
    return max(os.path.getmtime(p) for p in [os.path.join(CSRC, s) for s in SOURCES] + _header_paths())


def _compile(nvcc: str, src: str, obj: str, extra: list[str]) -> None:
    subprocess.run([nvcc, *NVCC_FLAGS_COMPILE, *extra, "-c", "-o", obj, src], check=True)


def _build(out: str, extra: list[str], tag: str, verbose: bool) -> str:
    """One nvcc -c per source file (in parallel: the tcgen05 kernels take a minute each), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(PKG_DIR, "build", tag)
    os.makedirs(objdir, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(p) for p in _header_paths())
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s[:-3] + ".o")
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_mtime):
            jobs.append((src, obj))
    if verbose and jobs:
        print(f"nvcc {' '.join(NVCC_FLAGS_COMPILE + extra)} -c  [{', '.join(os.path.basename(j[0]) for j in jobs)}]")
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
        for f in [ex.submit(_compile, nvcc, src, obj, extra) for src, obj in jobs]:
            f.result()
    objs = [os.path.join(objdir, s[:-3] + ".o") for s in SOURCES]
    subprocess.run([nvcc, *NVCC_FLAGS_LINK, "-o", out, *objs, "-ldl"], check=True)
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libsrk.so unless an up-to-date build exists. Returns the library path."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    if force:
        shutil.rmtree(os.path.join(PKG_DIR, "build", "release"), ignore_errors=True)
    return _build(LIB_PATH, [], "release", verbose)


def build_trace_library() -> str:
    """Development build with the in-kernel timeline enabled (tools/trace_conv.py).  Written under build/ (git- and
    gpurun-ignored unless a tool asks for it by path); never loaded by the product loader."""
    out = os.path.join(PKG_DIR, "build", "libsrk_trace.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    return _build(out, ["-DSRK_TRACE"], "trace", False)


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
