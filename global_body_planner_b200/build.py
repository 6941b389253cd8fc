"""Builds libgbp_b200.so (the C-ABI library, include/gbp_b200.h) in-tree with nvcc for sm_100a."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libgbp_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "gbp_capi.cu")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in ("gbp_device.cuh", "gbp_kernels.cuh", "gbp_planner.cuh")] + \
          [os.path.join(HERE, "..", "include", "gbp_b200.h")]
# -fmad=false: fp64 results must match the reference's x86-64 (no FMA) arithmetic bit for bit.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + SOURCES
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    import sys
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
