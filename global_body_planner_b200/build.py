"""Builds libgbp_b200.so (the C-ABI library, include/gbp_b200.h) in-tree with nvcc for sm_100a."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libgbp_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "gbp_capi.cu")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in ("gbp_device.cuh", "gbp_kernels.cuh", "gbp_planner.cuh", "gbp_walk.cuh")] + \
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


HOST_SO = os.path.join(HERE, "libglobal_body_planner_b200.so")
HOST_SOURCES = [os.path.join(HERE, "host", f) for f in ("dropin.cpp", "planners.cpp", "global_body_planner.cpp")]
CLI = os.path.join(HERE, "gbp_plan")
CLI_SOURCE = os.path.join(HERE, "host", "gbp_plan_main.cpp")
INCLUDE = os.path.join(HERE, "..", "include")


def build_host(force=False):
    """The drop-in C++ classes (include/global_body_planner/*.h) over the C ABI."""
    hdrs = [os.path.join(INCLUDE, "global_body_planner", f) for f in os.listdir(os.path.join(INCLUDE, "global_body_planner"))]
    if not force and os.path.exists(HOST_SO) and os.path.exists(CLI) and \
            all(os.path.getmtime(f) <= min(os.path.getmtime(HOST_SO), os.path.getmtime(CLI)) for f in HOST_SOURCES + hdrs + [SO, CLI_SOURCE]):
        return HOST_SO
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-std=c++14", "-O2", "-fPIC", "-shared", "-Wall", "-I" + INCLUDE, "-o", HOST_SO] + HOST_SOURCES + \
          ["-L" + HERE, "-lgbp_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    # gbp_plan: the ROS-free callPlanner driver as a command-line tool
    subprocess.run([cxx, "-std=c++14", "-O2", "-Wall", "-I" + INCLUDE, "-o", CLI, CLI_SOURCE, "-L" + HERE, "-lglobal_body_planner_b200",
                    "-lgbp_b200", "-Wl,-rpath,$ORIGIN"], check=True)
    return HOST_SO


if __name__ == "__main__":
    import sys
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_host(force="--force" in sys.argv)
    print(SO)
