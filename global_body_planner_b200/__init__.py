"""global_body_planner_b200 — B200-native RRT-Connect extend path behind the reference's API.

The product is the C-ABI library ``libgbp_b200.so`` (include/gbp_b200.h) plus the drop-in C++ classes in
include/global_body_planner/.  This Python package is only the thin ctypes binding the tests and
bench.py use to reach that C ABI; it holds no algorithmic code and has NO CPU fallback: importing
``capi`` without the built library, or calling it without a CUDA device, raises.
"""
from .capi import (ADVANCED, FLAG_NEAR, FLAG_OOG, FLAG_VALID, FLIGHT, FORWARD, REACHED, REVERSE, STANCE, TRAPPED,
                   GbpError, PlanParams, SvParams, SvResult, States, sv_params, pack_rows, unpack_bits, PLAN_STATS_DTYPE, Terrain, Tree, lib, propagate, sample_actions, valid_actions,
                   distance, version, device_count, set_device, interp_path, max_curvature, own_map_layer, rotate_grf, curvature)

__all__ = ["Terrain", "Tree", "States", "PlanParams", "SvParams", "SvResult", "sv_params", "pack_rows", "unpack_bits", "PLAN_STATS_DTYPE", "GbpError", "lib", "propagate", "sample_actions",
           "valid_actions", "distance", "interp_path", "max_curvature", "own_map_layer", "rotate_grf", "curvature", "version", "device_count", "set_device", "FORWARD", "REVERSE", "FLIGHT", "STANCE",
           "TRAPPED", "ADVANCED", "REACHED", "FLAG_VALID", "FLAG_OOG", "FLAG_NEAR"]
