"""Builds the sm_100a shared library in-tree (h264_fer_b200/libfh264_b200.so) with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot. nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "fh264_b200.cu")
LIB = os.path.join(HERE, "libfh264_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    srcs.append(os.path.join(HERE, "..", "include", "fh264_b200.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


LIB_BOUNDS = os.path.join(HERE, "libfh264_b200_bounds.so")


def build_bounds() -> str:
    """Debug build with -DFH_BOUNDS (index checks on the shared-memory work arrays of the search kernels); load it with
    FH264_B200_LIB=<path> to run the parity suite on it."""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DFH_BOUNDS", "-o", LIB_BOUNDS, SRC]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout.decode()[-4000:])
    return LIB_BOUNDS


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB, SRC]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        log = res.stdout.decode()
        with open(os.path.join(HERE, "build.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + log[-4000:])
        if verbose:
            print(log)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_bounds() if "--bounds" in sys.argv else build(force=True, verbose=True))
