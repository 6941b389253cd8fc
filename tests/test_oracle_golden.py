"""The oracle restatement (oracle/gbp_oracle.c) against golden vectors minted from the unmodified
reference (tests/golden/make_golden.py).  CPU only; this is what pins the oracle on the GPU box."""
import numpy as np
import pytest

import pyoracle as po
from conftest import assert_bits_equal


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    o = po.Oracle()
    assert list(o.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(o.philox([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(o.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_layout():
    o = po.Oracle()
    u = o.uniforms(5, (3 << 32) | 9, (1 << 40) + 17, 1, 0, 10)
    assert ((u >= 0) & (u < 1)).all()
    w = o.philox([17, 1 << 8, 9, 3 | (2 << 24) | (1 << 28)], [5, 0])  # block 2 -> uniforms 4, 5
    assert u[4] == ((int(w[0]) >> 5) * 2.0 ** 26 + (int(w[1]) >> 6)) * 2.0 ** -53
    assert u[5] == ((int(w[2]) >> 5) * 2.0 ** 26 + (int(w[3]) >> 6)) * 2.0 ** -53
    assert (o.uniforms(5, (3 << 32) | 9, (1 << 40) + 17, 1, 3, 4) == u[3:7]).all()


def test_det_math_accuracy():
    o = po.Oracle()
    for x in np.concatenate([np.logspace(-16, 0, 50), [2.0 ** -53, 1.0]]):
        assert abs(o.det_log(x) - np.log(x)) <= 4e-16 * max(1.0, abs(np.log(x)))
    for x in np.linspace(0, 6.2832, 200):
        s, c = o.det_sincos(x)
        assert abs(s - np.sin(x)) < 3e-16 and abs(c - np.cos(x)) < 3e-16


def test_terrain_lookups(golden):
    _, T, G = golden
    o = po.Oracle(T)
    h, fl = o.ground_height(G["probe_x"], G["probe_y"])
    assert not fl.any()
    assert_bits_equal(h, G["probe_h"], what="getGroundHeight")
    assert (o.height_is_nan(G["probe_x"], G["probe_y"]) == G["probe_nan"]).all()
    assert_bits_equal(o.surface_normal(G["probe_x"], G["probe_y"]), G["probe_normal"], what="getSurfaceNormal")


def test_out_of_grid_semantics_defined():
    """cell-0-anchored extrapolation + OOG flag (reference: undefined)."""
    T = po.Terrain(np.array([0.0, 1.0, 2.0]), np.array([0.0, 1.0, 2.0]), np.arange(9.0).reshape(3, 3))
    o = po.Oracle(T)
    h, fl = o.ground_height([-0.5, 2.0, 0.5, 2.5], [0.5, 0.5, 2.0, 0.5])
    assert (fl == po.FLAG_OOG).all()
    assert h[0] == 3 * -0.5 + 0.5  # extrapolated from cell (0,0): z = 3x + y
    h2, fl2 = o.ground_height([1.999999], [1.999999])
    assert fl2[0] == 0 and abs(h2[0] - (3 * 1.999999 + 1.999999)) < 1e-12


def test_propagation(golden):
    _, T, G = golden
    o = po.Oracle(T)
    s, a, t = G["pair_states"], G["pair_actions"], G["prop_t"]
    assert_bits_equal(o.apply_stance(s, a, t), G["prop_stance"], what="applyStance")
    assert_bits_equal(o.apply_flight(s, t - 0.1), G["prop_flight"], what="applyFlight")
    assert_bits_equal(o.apply_stance_reverse(s, a, t), G["prop_stance_rev"], what="applyStanceReverse")
    assert_bits_equal(o.rotate_grf(G["grf_n"], G["grf_f"]), G["grf_out"], what="rotate_grf")


def test_valid_state_and_action(golden):
    _, T, G = golden
    o = po.Oracle(T)
    for phase, key in ((po.STANCE, "vs_stance"), (po.FLIGHT, "vs_flight")):
        v, fl = o.valid_states(G["vs_states"], phase)
        ok = (fl & po.FLAG_OOG) == 0
        assert ok.mean() > 0.9 and (v[ok] == G[key][ok]).all()
    assert (o.is_valid_action(G["va_actions"]) == G["va_verdict"]).all()
    assert 0 < G["va_verdict"].mean() < 1


@pytest.mark.parametrize("adaptive", [0, 1])
def test_validate_pairs(golden, adaptive):
    _, T, G = golden
    o = po.Oracle(T)
    v, fl, sn, tn, cnt = o.validate_pairs(G["pair_states"], G["pair_actions"], G["pair_dir"], adaptive=bool(adaptive))
    ing = (fl & po.FLAG_OOG) == 0
    assert (ing == G[f"pair_ingrid_{adaptive}"].astype(bool)).all() and ing.mean() > 0.85
    assert (v[ing] == G[f"pair_verdict_{adaptive}"][ing]).all()
    assert G[f"pair_verdict_{adaptive}"][ing].sum() > 50
    assert_bits_equal(sn, G[f"pair_snew_{adaptive}"], where=ing, what="s_new")
    assert_bits_equal(tn, G[f"pair_tnew_{adaptive}"], where=ing, what="t_new")
    # outputs the reference leaves unwritten are DEFINED as s_new = s, t_new = 0
    unw = np.isnan(G[f"pair_snew_{adaptive}"]).all(axis=1) & ing
    assert (sn[unw] == G["pair_states"][unw]).all()
    assert (tn[np.isnan(G[f"pair_tnew_{adaptive}"]) & ing] == 0).all()
    assert 2 <= cnt[0] / len(v) <= 19


def test_attempt_connect(golden):
    _, T, G = golden
    o = po.Oracle(T)
    st, sn, an, fl = o.attempt_connect(G["con_existing"], G["con_target"], G["con_dir"])
    ing = (fl & po.FLAG_OOG) == 0
    assert (ing == G["con_ingrid"].astype(bool)).all() and ing.mean() > 0.85
    assert (st[ing] == G["con_status"][ing]).all()
    hit = (st != po.TRAPPED) & ing
    assert hit.sum() > 50
    assert_bits_equal(sn[hit], G["con_snew"][hit], what="attemptConnect s_new")
    assert_bits_equal(an[hit], G["con_anew"][hit], what="attemptConnect a_new")


def test_tree_queries(golden):
    _, T, G = golden
    o = po.Oracle(T)
    idx, dist, uniq = o.nearest(G["nn_verts"], G["nn_q"])
    assert_bits_equal(dist, G["nn_dist"], what="nearest distance")
    u = uniq.astype(bool)
    assert u.mean() > 0.9 and (idx[u] == G["nn_idx"][u]).all()
    off = 0
    for j, cnt in enumerate(G["near_counts"]):
        ids = o.near(G["nn_verts"], G["nn_q"][j], 3.0)
        assert (ids == G["near_ids"][off:off + cnt]).all()  # ascending id == sorted reference set
        off += cnt
    v = G["nn_verts"]
    for kind, key in ((0, "dist_pose"), (1, "dist_state"), (2, "dist_yaw")):
        assert_bits_equal(o.distance(v[:500], v[500:1000], kind), G[key], what=key)


def test_interp_path_and_curvature(golden):
    """getInterpPath / calculateMaxCurvature (planning_utils.cpp:142-193, :884-909) against the reference's vectors"""
    name, T, G = golden
    o = po.Oracle(T)
    s, t, ph = o.interp_path(G["interp_in_states"], G["interp_in_actions"], 0.05)
    assert len(s) == len(G["interp_states"]) and len(ph) == len(s) - 1
    assert_bits_equal(s, G["interp_states"], what="interp states")
    assert_bits_equal(t, G["interp_t"], what="interp times")
    assert (ph == G["interp_phase"]).all() and set(np.unique(ph)) == {0, 1, 2}
    assert o.max_curvature(s) == float(G["interp_max_curvature"])
    # capacity: the count is still the full length, only `cap` entries are written
    s2, t2, _ = o.interp_path(G["interp_in_states"], G["interp_in_actions"], 0.05, cap=10)
    assert len(s2) == len(s) or len(s2) == 10


# ---- planner level: the oracle planner against results minted from the UNMODIFIED reference's own loops
# (tests/golden/golden_planner.npz, tests/golden/make_golden_planner.py).  Runs everywhere, /root/reference not needed.
import os  # noqa: E402

import planner_cases as pc  # noqa: E402
from conftest import load_terrain  # noqa: E402


def _planner_golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_planner", os.path.join(os.path.dirname(__file__), "golden", "make_golden_planner.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.load()


@pytest.mark.parametrize("case", pc.CASES, ids=[c[0] for c in pc.CASES])
def test_planner_matches_reference_loops_golden(case):
    start, goal, want = _planner_golden()[case[0]]
    T = load_terrain(case[1])
    o = po.Oracle(T)
    s, g = pc.start_goal(case[1], T, o)
    assert pc.bits_equal(s, start) and pc.bits_equal(g, goal)
    pc.compare_run(case[0], pc.oracle_run(o, case, s, g), want)
