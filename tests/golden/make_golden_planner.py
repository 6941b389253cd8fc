#!/usr/bin/env python
"""Mints tests/golden/golden_planner.npz: the results of the UNMODIFIED reference's planner loops (runRRTConnect with the
real extend / newConfig / connect / RRT* extend, oracle/_ref/libgbp_ref_pin.so) on the cases of tests/planner_cases.py.
Needs /root/reference (build container only).  The GPU box has no reference: tests/test_oracle_golden.py replays the
file against the oracle planner and tests/test_gpu_planner.py against the device planner.

    python tests/golden/make_golden_planner.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import pyoracle as po  # noqa: E402
import planner_cases as pc  # noqa: E402


def main():
    po.build(ref=True)
    assert po.RefPin.available(), "needs /root/reference to build oracle/_ref/libgbp_ref_pin.so"
    out = {}
    for case in pc.CASES:
        T = po.Terrain.from_npz(os.path.join(HERE, f"terrain_{case[1]}.npz"))
        o, pin = po.Oracle(T), po.RefPin(T)
        s, g = pc.start_goal(case[1], T, o)
        r = pc.reference_run(pin, case, s, g)
        pc.compare_run(case[0], pc.oracle_run(o, case, s, g), r)  # refuse to mint a file the oracle disagrees with
        n = case[0]
        out[n + "/start"], out[n + "/goal"] = s, g
        out[n + "/scalars"] = np.array([r["solved"], r["iters"], r["nv_a"], r["nv_b"]], dtype=np.int64)
        out[n + "/path_stats"] = np.array([r["path_length"], r["path_yaw"], r["path_cost"]]) if r["solved"] else np.zeros(3)
        out[n + "/path_states"], out[n + "/path_actions"] = r["path_states"], r["path_actions"]
        for side in ("tree_a", "tree_b"):
            for k, v in r[side].items():
                out[f"{n}/{side}/{k}"] = v
        print(f"{n}: solved {r['solved']} after {r['iters']} iterations, {r['nv_a']} + {r['nv_b']} vertices, path {len(r['path_states'])} states")
        pin.close()
    np.savez_compressed(os.path.join(HERE, "golden_planner.npz"), **out)


def load(path=os.path.join(HERE, "golden_planner.npz")):
    """-> {case name: (start, goal, run dict in the layout planner_cases.compare_run expects)}"""
    d = np.load(path)
    res = {}
    for case in pc.CASES:
        n = case[0]
        sc = d[n + "/scalars"]
        run = dict(solved=int(sc[0]), iters=int(sc[1]), nv_a=int(sc[2]), nv_b=int(sc[3]), path_states=d[n + "/path_states"],
                   path_actions=d[n + "/path_actions"], path_length=float(d[n + "/path_stats"][0]), path_yaw=float(d[n + "/path_stats"][1]),
                   path_cost=float(d[n + "/path_stats"][2]))
        for side in ("tree_a", "tree_b"):
            run[side] = {k: d[f"{n}/{side}/{k}"] for k in ("states", "actions", "parent", "g", "yaw")}
        res[n] = (d[n + "/start"], d[n + "/goal"], run)
    return res


if __name__ == "__main__":
    main()
