"""The drop-in C++ API (include/global_body_planner/*.h, the reference's class names) driven by a
C++ caller (tests/cpp/test_dropin.cpp), checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import pyoracle as po
from conftest import ROOT, assert_bits_equal, load_terrain

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    import __graft_entry__ as entry
    entry.build()
    out = str(tmp_path_factory.mktemp("dropin") / "test_dropin")
    pkg = os.path.join(ROOT, "global_body_planner_b200")
    subprocess.run(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"),
                    "-o", out, "-L" + pkg, "-lglobal_body_planner_b200", "-lgbp_b200", "-Wl,-rpath," + pkg], check=True)
    return out


def run(exe, tmp_path, T, s, a, d, query=None):
    tb, cb, qb = str(tmp_path / "t.bin"), str(tmp_path / "c.bin"), str(tmp_path / "q.bin")
    with open(tb, "wb") as f:
        np.array([T.nx, T.ny], np.float64).tofile(f)
        for arr in (T.x, T.y, T.z, T.dx, T.dy, T.dz):
            np.ascontiguousarray(arr, np.float64).tofile(f)
    with open(cb, "wb") as f:
        np.array([len(s)], np.float64).tofile(f)
        for arr in (s, a, d.astype(np.float64)):
            np.ascontiguousarray(arr, np.float64).tofile(f)
    args = [exe, tb, cb]
    if query is not None:
        np.ascontiguousarray(np.concatenate(query), np.float64).tofile(qb)
        args.append(qb)
    txt = subprocess.run(args, check=True, capture_output=True, text=True, timeout=120).stdout
    out = {}
    for ln in txt.splitlines():
        k, *v = ln.split()
        if k in ("RRT", "Path", "Failed,"):
            continue
        out.setdefault(k, []).append(np.array([float(x) for x in v]))
    return out


@pytest.mark.parametrize("name", ["slope", "rough_terrain"])
def test_dropin_api_against_oracle(exe, tmp_path, name):
    T = load_terrain(name)
    G = np.load(os.path.join(ROOT, "tests", "golden", f"golden_{name}.npz"))
    o = po.Oracle(T)
    keep = G["pair_ingrid_0"].astype(bool)
    s, a, d = G["pair_states"][keep][:400], G["pair_actions"][keep][:400], G["pair_dir"][keep][:400]
    n = len(s)
    h = o.ground_height(s[:, 0], s[:, 1])[0]
    query = None
    if name == "slope":  # a query the iteration-budgeted oracle planner solves, so the anytime loop terminates
        from test_gpu_planner import queries
        qs, qg = queries(o, T, 24, 5)
        P = po.PlanParams(6, 0, 400, 256, 0, 0, 0)
        solvable = [i for i in range(len(qs)) if o.plan(qs[i], qg[i], 1, 1 << 20, P)[0].solved]
        assert solvable
        query = (qs[solvable[0]], qg[solvable[0]])
    R = run(exe, tmp_path, T, s, a, d, query)
    m = len(R["scalar"])
    sc = np.array(R["scalar"])
    assert_bits_equal(sc[:, 0], h[:m], what="getGroundHeight")
    assert_bits_equal(sc[:, 1:4], o.surface_normal(s[:m, 0], s[:m, 1]), what="getSurfaceNormal")
    assert (sc[:, 5] == o.valid_states(s[:m], po.STANCE)[0]).all() and (sc[:, 6] == o.valid_states(s[:m], po.FLIGHT)[0]).all()
    assert (sc[:, 7] == o.is_valid_action(a[:m])).all()
    nxt = s[(np.arange(m) + 1) % n]
    assert_bits_equal(sc[:, 8], o.distance(s[:m], nxt, 0), what="poseDistance")
    assert_bits_equal(sc[:, 9], o.distance(s[:m], nxt, 1), what="stateDistance")
    vo, fo, sno, tno, _ = o.validate_pairs(s, a, d)
    pr = np.array(R["pair"])
    assert (pr[:, 0] == vo[:m]).all()
    assert_bits_equal(pr[:, 1], tno[:m], what="t_new")
    assert_bits_equal(pr[:, 2:], sno[:m], what="s_new")
    assert_bits_equal(np.array(R["stance"]), o.apply_stance(s[:m], a[:m], 0.1), what="applyStance")
    assert_bits_equal(np.array(R["flight"]), o.apply_flight(s[:m], 0.2), what="applyFlight")
    assert_bits_equal(np.array(R["stancerev"]), o.apply_stance_reverse(s[:m], a[:m], 0.1), what="applyStanceReverse")
    assert (R["batchverdict"][0] == vo).all()
    nv = min(n, 200)
    qn = np.arange(nv, min(n, nv + 40))
    assert (R["nearest"][0] == o.nearest(s[:nv], s[qn])[0]).all()
    parent = np.array([-1] + [(i - 1) // 2 for i in range(1, nv)])
    g = np.zeros(nv)
    for i in range(1, nv):
        g[i] = g[parent[i]] + o.distance(s[parent[i]][None], s[i][None], 0)[0]
    assert_bits_equal(R["gvalues"][0], g, what="g values")
    assert (R["near"][0] == o.near(s[:nv], s[nv], 3.0)).all()
    assert tuple(R["treemisc"][0]) == (nv, 3, 2)
    idx = np.arange(0, min(n - 1, 60), 2)
    so, sno2, ano, flo = o.attempt_connect(s[idx], s[idx + 1], d[idx])
    cn = np.array(R["connect"])
    assert (cn[:, 0] == so).all()
    hit = so != po.TRAPPED
    assert_bits_equal(cn[hit, 1:9], sno2[hit], what="attemptConnect s_new")
    assert_bits_equal(cn[hit, 9:], ano[hit], what="attemptConnect a_new")
    ext = R["extend"][0]
    assert ext[1] == 1 + ext[0]
    for chk in R.get("extendcheck", []):
        assert chk[0] == 1 and chk[1] == 0 and chk[2] == 1  # appended vertices are the valid end states of their actions
    if name == "slope":
        p = R["plan"][0]
        assert p[0] >= 2 and p[1] == 1 and p[2] < 1e-9 and p[3] < 1e-9, f"buildRRTConnect path is not a valid chain: {p}"
        assert p[9] > p[0]


def write_csv_dir(T, directory):
    """the reference's data/<name>/ layout: rows = y, columns = x"""
    os.makedirs(directory, exist_ok=True)
    X, Y = np.meshgrid(T.x, T.y)
    for k, v in {"x": X, "y": Y, "z": T.z.T, "dx": T.dx.T, "dy": T.dy.T, "dz": T.dz.T}.items():
        with open(os.path.join(directory, f"{k}data.csv"), "w") as f:
            for row in v:
                f.write(",".join(repr(float(c)) for c in row) + "\n")


def test_call_planner_driver(tmp_path):
    """gbp_plan = GlobalBodyPlanner::callPlanner without ROS (global_body_planner.cpp:60-168) on BASELINE configs[0]:
    data/slope from its CSV files, (0,0) -> (8,0), body 0.30 m above ground (the fork's 0.375 m start pose is
    invalid on this map), the reference's statistics protocol; the discrete plan is re-validated primitive by primitive
    by the oracle and the published body plan must be the oracle's interpolation of it."""
    import __graft_entry__ as entry
    entry.build()
    T = load_terrain("slope")
    o = po.Oracle(T)
    d = str(tmp_path / "slope")
    write_csv_dir(T, d)
    exe = os.path.join(ROOT, "global_body_planner_b200", "gbp_plan")
    plan, disc = str(tmp_path / "plan.csv"), str(tmp_path / "disc.csv")
    r = subprocess.run([exe, d, "--height", "0.30", "--num-calls", "2", "--replan-time-limit", "0.3", "--max-time-solve", "40",
                        "--attempts", "2048", "1500", "512", "--plan-out", plan, "--discrete-out", disc],
                       capture_output=True, text=True, timeout=170)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "----- plan times: 2 / 2 -----" in out and "---------- Average ----------" in out and "Average path length:" in out
    summ = dict(zip(out.split("SUMMARY")[1].split()[0::2], out.split("SUMMARY")[1].split()[1::2]))
    assert int(summ["calls"]) == 2
    rows = np.loadtxt(disc, delimiter=",", ndmin=2)
    assert len(rows) >= 2, "the planner found no plan for the slope query within 40 s"
    ss, aa = rows[:, :8], rows[:-1, 8:]
    h0 = o.ground_height(np.array([0.0, 8.0]), np.array([0.0, 0.0]))[0]
    assert np.allclose(ss[0], [0, 0, 0.30 + h0[0], 1, 0, 0, 0, 0]) and np.allclose(ss[-1], [8, 0, 0.30 + h0[1], 1, 0, 0, 0, 0])
    # every primitive joins its two states, and passes the pair check in the direction it was grown in (tree A: FORWARD
    # from its start state; tree B: REVERSE from its end state, rrt_connect.cpp:230-314; the two sample grids differ)
    end = o.apply_flight(o.apply_stance(ss[:-1], aa, aa[:, 6]), aa[:, 7])
    assert np.abs(end - ss[1:]).max() < 1e-9
    vf = o.validate_pairs(ss[:-1], aa, np.zeros(len(aa), np.uint8))[0]
    vr = o.validate_pairs(ss[1:], aa, np.ones(len(aa), np.uint8))[0]
    assert ((vf == 1) | (vr == 1)).all(), "a primitive of the returned plan is invalid under the oracle in both directions"
    body = np.loadtxt(plan, delimiter=",", ndmin=2)
    ips, ipt, ipp = o.interp_path(ss, aa, 0.05)
    assert len(body) == len(ips)
    assert_bits_equal(body[:, 1:9], ips, what="body plan states")
    assert np.array_equal(body[:, 0], ipt) and np.array_equal(body[:-1, 9].astype(int), ipp) and body[-1, 9] == -1
    # path_length_ of postProcessPath only sums the shortcut segments (rrt_connect.cpp:139-227, kept as is); the
    # geometric length of the discrete plan cannot be below the straight line
    assert float(summ["avg_path_length"]) > 0
    assert o.distance(ss[:-1], ss[1:], 0).sum() >= 8.0 - 1e-9


def test_call_planner_driver_rrt_star(tmp_path):
    """the same driver with algorithm = rrt-star-connect (device-resident RRT*-Connect attempts, rrt_star_connect.cpp)"""
    import __graft_entry__ as entry
    entry.build()
    T = load_terrain("slope")
    o = po.Oracle(T)
    d = str(tmp_path / "slope")
    write_csv_dir(T, d)
    exe = os.path.join(ROOT, "global_body_planner_b200", "gbp_plan")
    disc = str(tmp_path / "disc.csv")
    r = subprocess.run([exe, d, "--height", "0.30", "--algorithm", "rrt-star-connect", "--num-calls", "1", "--replan-time-limit", "0.2",
                        "--max-time-solve", "40", "--discrete-out", disc, "--quiet"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = np.loadtxt(disc, delimiter=",", ndmin=2)
    assert len(rows) >= 2, "no RRT*-Connect plan for the slope query within 40 s"
    ss, aa = rows[:, :8], rows[:-1, 8:]
    end = o.apply_flight(o.apply_stance(ss[:-1], aa, aa[:, 6]), aa[:, 7])
    assert np.abs(end - ss[1:]).max() < 1e-9
    vf = o.validate_pairs(ss[:-1], aa, np.zeros(len(aa), np.uint8))[0]
    vr = o.validate_pairs(ss[1:], aa, np.ones(len(aa), np.uint8))[0]
    assert ((vf == 1) | (vr == 1)).all()
    bad = subprocess.run([exe, d, "--algorithm", "a-star"], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 1 and "Invalid algorithm specified" in bad.stderr
