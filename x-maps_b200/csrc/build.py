#!/usr/bin/env python
"""Build ``x-maps_b200/libxmaps_b200.so`` (the C-ABI library) for sm_100a with nvcc.

    python x-maps_b200/csrc/build.py [--force] [--verbose]

The library links the CUDA runtime statically and has no Python / torch dependency.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libxmaps_b200.so")
SOURCES = ["xm_capi.cu"]
DEPS = SOURCES + ["xm_device.cuh", "xm_frame_kernels.cuh", "xm_stage_kernels.cuh", "xm_fused_kernel.cuh", "xm_batch_kernel.cuh", "xm_stream_kernels.cuh", os.path.join("..", "..", "include", "xmaps_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-rdc=true",  # K1 tail-launches its fix-up from the device (CUDA dynamic parallelism)
    "-Xcompiler", "-fPIC", "-shared",
]
LINK_LIBS = ["-lcudadevrt"]


STAMP = OUT + ".srchash"


def source_hash():
    """sha256 over every source the library is built from + the flags: the built library carries the hash of
    its sources next to it, so a stale .so (e.g. one that travelled with a snapshot whose sources changed, or
    whose mtimes were reset by a copy) is rebuilt, and an up-to-date one is not."""
    import hashlib

    h = hashlib.sha256()
    for d in sorted(DEPS):
        with open(os.path.join(HERE, d), "rb") as fh:
            h.update(d.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS + LINK_LIBS + os.environ.get("XM_NVCC_EXTRA", "").split()).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(OUT) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    extra = os.environ.get("XM_NVCC_EXTRA", "").split()  # e.g. -DXM_ROW_PHASE_A=0 for A/B builds
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES + LINK_LIBS
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if verbose or proc.returncode:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode:
        raise RuntimeError("nvcc failed building libxmaps_b200.so")
    with open(STAMP, "w") as fh:
        fh.write(source_hash() + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
