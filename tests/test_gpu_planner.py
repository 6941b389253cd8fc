"""Tier-2 parity: the device-resident batch planner (gbp_plan_batch, one warp per query) against the
iteration-budgeted oracle planner on the same Philox stream — trees, statistics and paths."""
import numpy as np
import pytest

import pyoracle as po
from conftest import assert_bits_equal, load_terrain

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gbp():
    import __graft_entry__ as entry
    entry.build()
    import global_body_planner_b200 as g
    assert g.device_count() > 0
    return g


def queries(o, T, n, seed):
    q = o.sample_states(seed, 1, 0, 40 * n)
    q[:, 3:8] = 0; q[:, 3] = 0.5
    v, _ = o.valid_states(q, po.STANCE)
    q = q[v == 1]
    s, g = [], []
    for i in range(len(q)):
        d = np.hypot(q[:, 0] - q[i, 0], q[:, 1] - q[i, 1])
        j = np.nonzero((d > 1.0) & (d < 3.0))[0]
        if len(j):
            s.append(q[i]); g.append(q[j[0]])
        if len(s) == n:
            break
    return np.array(s), np.array(g)


@pytest.mark.parametrize("name,K,best,star,post", [("slope", 6, 0, 0, 0), ("rough_terrain", 6, 0, 0, 0), ("slope", 48, 1, 0, 0),
                                                  ("synth_nan", 6, 0, 0, 0), ("slope", 6, 0, 0, 1), ("synth_nan", 24, 1, 0, 1),
                                                  ("slope", 32, 1, 1, 0), ("synth_nan", 32, 1, 1, 1), ("rough_terrain", 32, 1, 1, 0),
                                                  # fp32 NaN-free map on uniform axes: the mixed-precision / texture-gather evaluator
                                                  ("synth_mixed", 6, 0, 0, 0), ("synth_mixed", 24, 1, 0, 1), ("synth_mixed", 32, 1, 1, 1),
                                                  # closest-valid over more than a warp of candidates (several batches of 32 per newConfig)
                                                  ("rough_terrain", 200, 1, 0, 0), ("synth_mixed", 100, 1, 0, 1), ("slope", 70, 1, 1, 0)])
def test_plan_batch_matches_oracle(gbp, name, K, best, star, post):
    """RRT-Connect, RRT*-Connect (choose parent + rewire) and postProcessPath, all resident on the device."""
    T = load_terrain(name)
    o = po.Oracle(T)
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    s, g = queries(o, T, 24, 5)
    assert len(s) >= 8
    iters = 250 if star else 400
    P = gbp.PlanParams(K, best, iters, 256, 0, star, post)
    st, ps, pa = t.plan_batch(s, g, 9, 100, P, path_cap=128)
    Po = po.PlanParams(K, best, iters, 256, 0, star, post)
    nsolved = 0
    for i in range(len(s)):
        so, pso, pao = o.plan(s[i], g[i], 9, 100 + i, Po)
        for key in ("solved", "iters", "nv_a", "nv_b", "path_states", "pair_checks", "nn_queries"):
            assert st[key][i] == getattr(so, key), f"query {i}: {key}"
        if so.solved:
            nsolved += 1
            n = so.path_states
            assert_bits_equal(np.array([st["path_length"][i]]), np.array([so.path_length]), what="path length")
            assert abs(st["path_yaw"][i] - so.path_yaw) < 1e-9  # yaw sums go through atan2 (libm): tolerance, not bits
            assert_bits_equal(np.array([st["path_duration"][i]]), np.array([so.path_duration]), what="path duration")
            assert_bits_equal(ps[i, :n], pso, what="path states")
            assert_bits_equal(pa[i, :n - 1], pao, what="path actions")
    assert nsolved >= 1  # path stitching / path statistics assertions ran on every map, data/rough_terrain included


# ---- against the UNMODIFIED reference's own loops (tests/golden/golden_planner.npz, minted by
# tests/golden/make_golden_planner.py from oracle/_ref/libgbp_ref_pin.so: runRRTConnect with the real extend / newConfig /
# connect / RRT* extend on the shared Philox stream)
import planner_cases as pc  # noqa: E402


def _planner_golden():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden_planner", os.path.join(os.path.dirname(__file__), "golden", "make_golden_planner.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.load()


def _gpu_params(gbp, case, post_process=0):
    _, _, star, iters, adaptive, opts, _, _ = case
    return gbp.PlanParams(6, 0, iters, pc.CAP, adaptive, star, post_process, 0, opts[0], opts[1], opts[2], opts[3], pc.THRESH, pc.THRESH,
                          pc.W_LEN, pc.W_YAW)


@pytest.mark.parametrize("case", pc.CASES, ids=[c[0] for c in pc.CASES])
def test_plan_batch_matches_reference_loops(gbp, case):
    """The device planner against the reference's runRRTConnect: both trees vertex by vertex (states, actions, parents,
    g bit for bit; yaw sums to 1e-9: the device's atan2 is not glibc's), iterations, the stitched path and its length /
    yaw / cost — RRT-Connect and RRT*-Connect, fixed and adaptive step, with and without the fork's options (directional
    state / action sampling, yaw-aware cost), including data/slope and data/rough_terrain run to their first solution."""
    start, goal, want = _planner_golden()[case[0]]
    T = load_terrain(case[1])
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    st, ps, pa, trees = t.plan_batch_trees(start[None], goal[None], case[6], case[7], _gpu_params(gbp, case), path_cap=256, tree_cap=pc.CAP)
    n = int(st["path_states"][0])
    got = dict(solved=int(st["solved"][0]), iters=int(st["iters"][0]), nv_a=int(st["nv_a"][0]), nv_b=int(st["nv_b"][0]), tree_a=trees[0][0],
               tree_b=trees[0][1], path_states=ps[0, :n], path_actions=pa[0, :max(n - 1, 0)], path_length=float(st["path_length"][0]),
               path_yaw=float(st["path_yaw"][0]), path_cost=float(st["path_cost"][0]))
    pc.compare_run(case[0], got, want, yaw_tol=1e-9)


def test_fork_options_change_the_search(gbp):
    """the options are not silently ignored on the batch path: switching one on changes the trees"""
    case = [c for c in pc.CASES if c[0] == "rough_short"][0]
    start, goal, _ = _planner_golden()[case[0]]
    T = load_terrain(case[1])
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    base = t.plan_batch(start[None], goal[None], 7, 0, gbp.PlanParams(6, 0, 3000, 256, 0, 0, 0, 0))
    for opts in ((1, 0, 0, 0), (0, 0, 1, 0)):
        P = gbp.PlanParams(6, 0, 3000, 256, 0, 0, 0, 0, opts[0], opts[1], opts[2], opts[3], 0.5, 0.5, 1.0, 1.0)
        other = t.plan_batch(start[None], goal[None], 7, 0, P)
        assert (other["nv_a"][0], other["nv_b"][0], other["pair_checks"][0]) != (base["nv_a"][0], base["nv_b"][0], base["pair_checks"][0])


def test_plan_batch_capacity_and_budget(gbp):
    T = load_terrain("slope"); o = po.Oracle(T)
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    s, g = queries(o, T, 8, 6)
    P = gbp.PlanParams(6, 0, 300, 4, 0, 0, 0)  # tiny trees: queries stop when a tree is full
    st = t.plan_batch(s, g, 3, 0, P)
    Po = po.PlanParams(6, 0, 300, 4, 0, 0, 0)
    for i in range(len(s)):
        so, _, _ = o.plan(s[i], g[i], 3, i, Po)
        assert (st["solved"][i], st["iters"][i], st["nv_a"][i], st["nv_b"][i]) == (so.solved, so.iters, so.nv_a, so.nv_b)
    assert (st["nv_a"] <= 4).all() and (st["nv_b"] <= 4).all()


@pytest.mark.parametrize("mode", ["mega", "pipe"])
def test_stop_after_solved(gbp, monkeypatch, mode):
    """anytime use (many attempts at one query): the launch ends once `stop_after_solved` attempts have solved; attempts
    that did solve are unaffected (their own search is deterministic), the others report the work done so far.  Both forms
    of the batch planner (the pipelined one counts solved queries in k_pipe_connect and stops the others in k_pipe_prep)."""
    monkeypatch.setenv("GBP_PLAN_MODE", mode)
    T = load_terrain("slope"); o = po.Oracle(T)
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    s, g = queries(o, T, 24, 5)
    Po = po.PlanParams(6, 0, 400, 256, 0, 0, 0)
    i = next(k for k in range(len(s)) if o.plan(s[k], g[k], 1, 1 << 20, Po)[0].solved)
    n = 4096
    S, G = np.repeat(s[i][None], n, 0), np.repeat(g[i][None], n, 0)
    full = t.plan_batch(S, G, 1, 1 << 20, gbp.PlanParams(6, 0, 400, 256, 0, 0, 0, 0))
    assert t.plan_batch_form(gbp.PlanParams(6, 0, 400, 256, 0, 0, 0, 4), n) == {"mega": "megakernel", "pipe": "pipelined"}[mode]
    early = t.plan_batch(S, G, 1, 1 << 20, gbp.PlanParams(6, 0, 400, 256, 0, 0, 0, 4))
    assert full["solved"][0] == 1 and full["solved"].sum() >= 4
    assert 4 <= early["solved"].sum() <= full["solved"].sum()
    hit = early["solved"] == 1
    for key in ("iters", "nv_a", "nv_b", "path_states", "pair_checks", "nn_queries", "path_length"):
        assert (early[key][hit] == full[key][hit]).all(), key
    assert (early["iters"] <= full["iters"]).all() and early["iters"].sum() < full["iters"].sum()


def test_pipelined_form_is_identical(gbp, monkeypatch):
    """GBP_PLAN_MODE=pipe: rounds of k_pipe_prep / k_walk_seg / k_pipe_triage / k_pipe_select / k_pipe_connect over all queries
    (the form large batches take) against the megakernel — statistics, paths and both trees of every query bit for bit,
    first-valid and closest-valid selection, directional action sampling, tiny trees (capacity stop) and tiny budgets."""
    T = load_terrain("synth_mixed"); o = po.Oracle(T)
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    s, g = queries(o, T, 24, 5)
    S, G = np.tile(s, (40, 1)), np.tile(g, (40, 1))
    for P in (gbp.PlanParams(6, 0, 300, 128, 0, 0, 1), gbp.PlanParams(12, 1, 150, 64, 0, 0, 0),
              gbp.PlanParams(6, 0, 200, 128, 0, 0, 0, 0, 0, 0, 1, 0, 0.0, 0.4, 1.0, 1.0), gbp.PlanParams(6, 0, 400, 6, 0, 0, 0),
              gbp.PlanParams(6, 0, 3, 128, 0, 0, 0)):
        monkeypatch.setenv("GBP_PLAN_MODE", "mega")
        assert t.plan_batch_form(P, len(S)) == "megakernel"
        a, pa_s, pa_a, ta = t.plan_batch_trees(S, G, 4, 50, P, path_cap=64, tree_cap=128)
        monkeypatch.setenv("GBP_PLAN_MODE", "pipe")
        assert t.plan_batch_form(P, len(S)) == "pipelined"
        # half-iterations a query speculates per round (k_pipe_prep) and the number of running queries at which the rest of the
        # batch moves to one warp per query (k_pipe_resume; 0: never, 1000000: at the first look): the results do not depend on them
        for spec, resume in (("8", None), ("1", None), ("2", "0"), ("4", "1000000"), ("16", "40")):
            monkeypatch.setenv("GBP_PIPE_SPEC", spec)
            if resume is None:
                monkeypatch.delenv("GBP_PIPE_RESUME", raising=False)
            else:
                monkeypatch.setenv("GBP_PIPE_RESUME", resume)
            b, pb_s, pb_a, tb = t.plan_batch_trees(S, G, 4, 50, P, path_cap=64, tree_cap=128)
            for k in a.dtype.names:
                assert np.array_equal(a[k], b[k]), (k, spec)
            for i in range(len(S)):
                n = int(a["path_states"][i])
                assert np.array_equal(pa_s[i, :n], pb_s[i, :n]) and np.array_equal(pa_a[i, :max(n - 1, 0)], pb_a[i, :max(n - 1, 0)])
            for (xa, xb), (ya, yb) in zip(ta, tb):
                for k in xa:
                    assert np.array_equal(xa[k], ya[k]) and np.array_equal(xb[k], yb[k]), (k, spec)
        monkeypatch.delenv("GBP_PIPE_SPEC")
        monkeypatch.delenv("GBP_PIPE_RESUME", raising=False)
    # the batch split into 3 groups of queries, each an independent pipeline on its own host thread and streams; and run as
    # consecutive sub-batches of 300 queries (what a batch larger than the device can hold does)
    monkeypatch.setenv("GBP_PIPE_GROUPS", "3")
    P = gbp.PlanParams(6, 0, 300, 128, 0, 0, 1)
    c, pc_s, pc_a, tc = t.plan_batch_trees(S, G, 4, 50, P, path_cap=64, tree_cap=128)
    monkeypatch.setenv("GBP_PIPE_GROUPS", "1")
    monkeypatch.setenv("GBP_PIPE_CHUNK", "300")
    d, pd_s, pd_a, td = t.plan_batch_trees(S, G, 4, 50, P, path_cap=64, tree_cap=128)
    monkeypatch.delenv("GBP_PIPE_GROUPS")
    monkeypatch.delenv("GBP_PIPE_CHUNK")
    assert c["solved"].sum() >= 5
    for k in c.dtype.names:
        assert np.array_equal(c[k], d[k]), k
    for i in range(len(S)):
        n = int(c["path_states"][i])
        assert np.array_equal(pc_s[i, :n], pd_s[i, :n]) and np.array_equal(pc_a[i, :max(n - 1, 0)], pd_a[i, :max(n - 1, 0)])
    for (xa, xb), (ya, yb) in zip(tc, td):
        for k in xa:
            assert np.array_equal(xa[k], ya[k]) and np.array_equal(xb[k], yb[k]), k
    monkeypatch.delenv("GBP_PLAN_MODE")
    assert t.plan_batch_form(gbp.PlanParams(6, 0, 300, 128, 0, 0, 0), 65536) == "pipelined"
    assert t.plan_batch_form(gbp.PlanParams(6, 0, 300, 128, 0, 0, 0), 1000) == "megakernel"
    assert t.plan_batch_form(gbp.PlanParams(6, 0, 300, 128, 0, 1, 0), 65536) == "megakernel"  # RRT* stays on the megakernel


def test_stepped_form_is_identical(gbp, monkeypatch):
    """GBP_PLAN_MODE=step: one launch per half-iteration over the whole batch (k_step_half) instead of the megakernel — same
    trees, statistics and paths, bit for bit"""
    T = load_terrain("synth_mixed"); o = po.Oracle(T)
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    s, g = queries(o, T, 24, 5)
    S, G = np.tile(s, (20, 1)), np.tile(g, (20, 1))
    P = gbp.PlanParams(6, 0, 300, 128, 0, 0, 1)
    monkeypatch.delenv("GBP_PLAN_MODE", raising=False)
    a, pa_s, pa_a, ta = t.plan_batch_trees(S, G, 4, 50, P, path_cap=64, tree_cap=128)
    monkeypatch.setenv("GBP_PLAN_MODE", "step")
    b, pb_s, pb_a, tb = t.plan_batch_trees(S, G, 4, 50, P, path_cap=64, tree_cap=128)
    assert a["solved"].sum() >= 5
    for k in a.dtype.names:
        assert np.array_equal(a[k], b[k]), k
    for i in range(len(S)):  # rows beyond a path's length are not written
        n = int(a["path_states"][i])
        assert np.array_equal(pa_s[i, :n], pb_s[i, :n]) and np.array_equal(pa_a[i, :max(n - 1, 0)], pb_a[i, :max(n - 1, 0)])
    for (xa, xb), (ya, yb) in zip(ta, tb):
        for k in xa:
            assert np.array_equal(xa[k], ya[k]) and np.array_equal(xb[k], yb[k]), k


@pytest.mark.parametrize("name", ["synth_mixed", "rough_terrain", "synth_nan"])
def test_device_wide_form_is_identical(gbp, monkeypatch, name):
    """GBP_PLAN_MODE=wide: one search at a time on a cooperative grid (k_plan_wide: newConfig's candidates spread over all SMs,
    one grid barrier per extend, every CTA keeps its own copy of the trees) against the megakernel — statistics, paths and
    both trees of every query bit for bit: K = 6 first-valid (32 lanes per candidate), closest-valid over 12 / 200 / 2000 / 4096 /
    8000 / 40000 candidates (32, 16, 8, 4, 1 lanes per candidate; survivors of the first lanes redistributed over the CTA; several passes), post-processing, the fork's directional samplers, a tree
    capacity stop, a tiny budget, the anytime stop."""
    T = load_terrain(name); o = po.Oracle(T)
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    S, G = queries(o, T, 6, 5)
    plain = None
    for P in (gbp.PlanParams(6, 0, 300, 128, 0, 0, 1), gbp.PlanParams(12, 1, 150, 64, 0, 0, 0), gbp.PlanParams(200, 1, 60, 64, 0, 0, 1),
              gbp.PlanParams(4096, 1, 12, 64, 0, 0, 0), gbp.PlanParams(40000, 1, 3, 64, 0, 0, 0), gbp.PlanParams(2000, 1, 12, 64, 0, 0, 0),
              gbp.PlanParams(8000, 1, 8, 64, 0, 0, 0), gbp.PlanParams(5000, 0, 8, 64, 0, 0, 0),
              gbp.PlanParams(6, 0, 200, 128, 0, 0, 0, 0, 1, 1, 1, 1, 0.4, 0.4, 1.0, 0.5), gbp.PlanParams(6, 0, 400, 6, 0, 0, 0),
              gbp.PlanParams(6, 0, 3, 128, 0, 0, 0), gbp.PlanParams(6, 0, 300, 128, 0, 0, 0, 2)):
        monkeypatch.setenv("GBP_PLAN_MODE", "mega")
        assert t.plan_batch_form(P, len(S)) == "megakernel"
        a, pa_s, pa_a, ta = t.plan_batch_trees(S, G, 4, 50, P, path_cap=64, tree_cap=128)
        monkeypatch.setenv("GBP_PLAN_MODE", "wide")
        assert t.plan_batch_form(P, len(S)) == "device-wide"
        b, pb_s, pb_a, tb = t.plan_batch_trees(S, G, 4, 50, P, path_cap=64, tree_cap=128)
        if P.stop_after_solved:  # anytime stop: queries run in order here, so the first `stop_after_solved` that solve are kept
            want = np.nonzero(plain["solved"])[0][:P.stop_after_solved]
            assert np.array_equal(np.nonzero(b["solved"])[0], want)
            last = want[-1] if len(want) == P.stop_after_solved else len(S) - 1
            assert np.array_equal(b["iters"][:last + 1], plain["iters"][:last + 1]) and not b["iters"][last + 1:].any()
            continue
        if plain is None:
            plain = a
        for k in a.dtype.names:
            assert np.array_equal(a[k], b[k]), (k, P.k_candidates)
        for i in range(len(S)):
            n = int(a["path_states"][i])
            assert np.array_equal(pa_s[i, :n], pb_s[i, :n]) and np.array_equal(pa_a[i, :max(n - 1, 0)], pb_a[i, :max(n - 1, 0)])
        for (xa, xb), (ya, yb) in zip(ta, tb):
            for k in xa:
                assert np.array_equal(xa[k], ya[k]) and np.array_equal(xb[k], yb[k]), k
    monkeypatch.delenv("GBP_PLAN_MODE")
    assert t.plan_batch_form(gbp.PlanParams(4096, 1, 300, 128, 0, 0, 0), 1) == "device-wide"   # a newConfig wider than the batch
    assert t.plan_batch_form(gbp.PlanParams(6, 0, 300, 128, 0, 0, 0), 1) == "megakernel"
    assert t.plan_batch_form(gbp.PlanParams(4096, 1, 300, 128, 0, 0, 0), 2368) == "megakernel"
    assert t.plan_batch_form(gbp.PlanParams(4096, 1, 300, 128, 0, 1, 0), 1) == "megakernel"    # RRT* stays on the megakernel
