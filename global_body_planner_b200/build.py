"""Builds libgbp_b200.so (the C-ABI library, include/gbp_b200.h) in-tree with nvcc for sm_100a."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libgbp_b200.so")
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")  # git-ignored and gpurun-ignored: only the linked .so travels
PUBLIC_H = os.path.join(HERE, "..", "include", "gbp_b200.h")
_COMMON = ["gbp_device.cuh", "gbp_kernels.cuh", "gbp_host.h"]
# translation unit -> private headers it includes (the units compile in parallel)
UNITS = {"gbp_capi.cu": _COMMON,
         "gbp_capi_validate.cu": _COMMON + ["gbp_walk.cuh", "gbp_sv.cuh"],
         "gbp_capi_plan.cu": _COMMON + ["gbp_planner.cuh"],
         "gbp_capi_pipeline.cu": _COMMON + ["gbp_planner.cuh", "gbp_walk.cuh", "gbp_sv.cuh", "gbp_pipeline.cuh"],
         "gbp_capi_wide.cu": _COMMON + ["gbp_planner.cuh", "gbp_wide.cuh"]}
SOURCES = [os.path.join(CSRC, u) for u in UNITS]
HEADERS = sorted({os.path.join(CSRC, h) for hs in UNITS.values() for h in hs}) + [PUBLIC_H]
# -fmad=false: fp64 results must match the reference's x86-64 (no FMA) arithmetic bit for bit.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in deps)


def stale():
    return _newer(SO, SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    jobs, objs = [], []
    for unit, hdrs in UNITS.items():
        src, obj = os.path.join(CSRC, unit), os.path.join(OBJ, unit[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src, PUBLIC_H] + [os.path.join(CSRC, h) for h in hdrs]):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            jobs.append((unit, subprocess.Popen(cmd)))
    failed = [unit for unit, p in jobs if p.wait() != 0]
    if failed:
        raise subprocess.CalledProcessError(1, "nvcc " + " ".join(failed))
    subprocess.run([nvcc, "-shared", "-o", SO] + objs, check=True)
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
