"""Planner-level parity cases shared by the live pin (tests/test_oracle_vs_ref.py::test_planner_loops_*), the golden
minting script (tests/golden/make_golden_planner.py) and the tests that replay the golden file on the GPU box
(tests/test_oracle_golden.py, tests/test_gpu_planner.py).

A case = one planning query run for a fixed number of runRRTConnect iterations on the Philox stream (seed, query):
terrain, start / goal, planner (RRT-Connect or RRT*-Connect), pair-check step mode and the fork's options
(state direction sampling [+ speed direction], action direction sampling, yaw-aware cost; thresholds 0.3, weights 1.0 / 0.5).
The long cases are the BASELINE queries run to their first solution: data/slope (0,0)->(8,0) at 0.30 m and data/rough_terrain
(0,0)->(8,0) at 0.375 m."""
import numpy as np

import pyoracle as po

THRESH, W_LEN, W_YAW = 0.3, 1.0, 0.5
#        name              terrain          star iters   adaptive opts(state, speed, action, yaw) seed query
CASES = [("slope_solve",    "slope",         0,   300000, 0,       (0, 0, 0, 0),                    7,   1),
         ("rough_solve",    "rough_terrain", 0,   330000, 0,       (0, 0, 0, 0),                    7,   1),
         ("rough_short",    "rough_terrain", 0,   20000,  0,       (0, 0, 0, 0),                    7,   0),
         ("slope_star",     "slope",         1,   60000,  0,       (0, 0, 0, 0),                    7,   0),
         ("rough_star",     "rough_terrain", 1,   60000,  0,       (0, 0, 0, 0),                    7,   1),
         ("rough_fork",     "rough_terrain", 0,   100000, 0,       (1, 1, 1, 1),                    7,   1),
         ("nan_fork",       "synth_nan",     0,   100000, 0,       (1, 0, 1, 1),                    7,   0),
         ("nan_star_yaw",   "synth_nan",     1,   30000,  0,       (0, 0, 0, 1),                    7,   1),
         ("mixed_star",     "synth_mixed",   1,   30000,  0,       (1, 1, 1, 0),                    7,   0),
         ("mixed_plain",    "synth_mixed",   0,   30000,  0,       (0, 0, 0, 0),                    7,   1),
         ("mixed_adaptive", "synth_mixed",   0,   30000,  1,       (0, 0, 1, 0),                    7,   2),
         ("slope_adaptive", "slope",         1,   20000,  1,       (1, 0, 0, 1),                    7,   3)]
CAP = 1024  # per-tree vertex capacity of every run (no case comes near it)


def start_goal(name, T, o):
    """the BASELINE query on the shipped maps; a 5 m traverse along the middle of the synthetic ones"""
    if name in ("rough_terrain", "slope"):
        x0, x1, y = 0.0, 8.0, 0.0
        body = 0.375 if name == "rough_terrain" else 0.30
    else:
        x0, x1, y, body = T.x[0] + 2.0, T.x[0] + 7.0, 0.5 * (T.y[0] + T.y[-1]), 0.30
    h = o.ground_height([x0, x1], [y, y])[0]
    return np.array([x0, y, h[0] + body, 1, 0, 0, 0, 0.0]), np.array([x1, y, h[1] + body, 1, 0, 0, 0, 0.0])


def oracle_params(case, post_process=0):
    _, _, star, iters, adaptive, opts, _, _ = case
    return po.PlanParams(6, 0, iters, CAP, adaptive, star, post_process, opts[0], opts[1], opts[2], opts[3], THRESH, THRESH, W_LEN, W_YAW)


def pin_params(case, sort_near=1):
    _, _, star, iters, adaptive, opts, _, _ = case
    return po.PinParams(star, iters, adaptive, sort_near, opts[0], opts[1], opts[2], opts[3], THRESH, THRESH, W_LEN, W_YAW)


def bits_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype == np.float64:
        return bool(np.array_equal(a.view(np.uint64), b.view(np.uint64)))
    return bool(np.array_equal(a, b))


def compare_run(tag, got, want, yaw_tol=0.0):
    """got / want: dict(solved, iters, nv_a, nv_b, tree_a, tree_b, path_states, path_actions, path_length, path_yaw, path_cost).
    Everything bit for bit; yaw sums within yaw_tol when a side computes atan2 with another libm (the GPU)."""
    for k in ("solved", "iters", "nv_a", "nv_b"):
        assert int(got[k]) == int(want[k]), f"{tag}: {k} {got[k]} != {want[k]}"
    for side in ("tree_a", "tree_b"):
        for k in ("states", "actions", "parent", "g"):
            assert bits_equal(got[side][k], want[side][k]), f"{tag}: {side}.{k} differs"
        if yaw_tol == 0.0:
            assert bits_equal(got[side]["yaw"], want[side]["yaw"]), f"{tag}: {side}.yaw differs"
        else:
            assert np.allclose(got[side]["yaw"], want[side]["yaw"], rtol=0, atol=yaw_tol), f"{tag}: {side}.yaw differs"
    if want["solved"]:
        assert bits_equal(got["path_states"], want["path_states"]), f"{tag}: path states differ"
        assert bits_equal(got["path_actions"], want["path_actions"]), f"{tag}: path actions differ"
        assert bits_equal(np.float64(got["path_length"]), np.float64(want["path_length"])), f"{tag}: path length"
        for k in ("path_yaw", "path_cost"):
            assert abs(got[k] - want[k]) <= yaw_tol, f"{tag}: {k} {got[k]} != {want[k]}"


def oracle_run(o, case, start, goal):
    st, ps, pa, ta, tb = o.plan_ex(start, goal, case[6], case[7], oracle_params(case), path_cap=4 * CAP)
    return dict(solved=st.solved, iters=st.iters, nv_a=st.nv_a, nv_b=st.nv_b, tree_a=ta, tree_b=tb, path_states=ps, path_actions=pa,
                path_length=st.path_length, path_yaw=st.path_yaw, path_cost=st.path_cost)


def reference_run(pin, case, start, goal, sort_near=1):
    r, ra, rb, ps, pa = pin.run(start, goal, case[6], case[7], pin_params(case, sort_near), cap=CAP, path_cap=4 * CAP)
    return dict(solved=r.solved, iters=(r.cells_used + 1) // 2, nv_a=r.nv_a, nv_b=r.nv_b, tree_a=ra, tree_b=rb, path_states=ps, path_actions=pa,
                path_length=r.path_length, path_yaw=r.path_yaw, path_cost=r.path_cost,
                oog_lookups=r.oog_height + r.oog_nan, near_sets=r.near_sets, near_reordered=r.near_reordered)
