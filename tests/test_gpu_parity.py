"""Parity tests proper: the CUDA path, called through the C ABI (libgbp_b200.so), against
(1) the golden vectors minted from the unmodified reference and (2) the oracle restatement on
freshly drawn batches.  Integer / index / verdict work must be bit-exact; fp64 outputs that do not
pass through libm (propagated states, t_new, heights, distances) must be bit-exact as well; the
yaw distance (atan2) is compared to 1e-12."""
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
    assert g.device_count() > 0, "no CUDA device: the GPU tests must not pass on a fallback"
    return g


@pytest.fixture(scope="module")
def dev(gbp, golden):
    name, T, G = golden
    return gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz), po.Oracle(T), T, G, name


def test_terrain_storage_choice(gbp):
    """fp32 cells iff lossless (shipped maps: 1 decimal -> fp64; synthetic fp32-rounded -> fp32)."""
    T = load_terrain("rough_terrain")
    assert gbp.Terrain(T.x, T.y, T.z).cell_bytes == 8
    T = load_terrain("synth_nan")
    assert gbp.Terrain(T.x, T.y, T.z).cell_bytes == 4


def test_terrain_lookups_golden(dev):
    t, o, T, G, _ = dev
    h, fl = t.ground_height(G["probe_x"], G["probe_y"])
    assert not fl.any()
    assert_bits_equal(h, G["probe_h"], what="getGroundHeight")
    assert (t.height_is_nan(G["probe_x"], G["probe_y"]) == G["probe_nan"]).all()
    assert_bits_equal(t.surface_normal(G["probe_x"], G["probe_y"]), G["probe_normal"], what="getSurfaceNormal")
    x, y = t.axes()
    assert (x == T.x).all() and (y == T.y).all()


def test_out_of_grid_matches_oracle(dev):
    t, o, T, G, _ = dev
    rng = np.random.default_rng(0)
    x = rng.uniform(T.x[0] - 1, T.x[-1] + 1, 4000); y = rng.uniform(T.y[0] - 1, T.y[-1] + 1, 4000)
    x[:4] = [T.x[-1], T.x[0], np.nan, T.x[3]]; y[:4] = [T.y[2], T.y[-1], T.y[2], np.inf]
    hg, fg = t.ground_height(x, y)
    ho, fo = o.ground_height(x, y)
    assert (fg == fo).all() and fo.any()
    assert_bits_equal(hg, np.where(np.isnan(ho), 0, ho), where=~np.isnan(ho), what="OOG height")


def test_propagation_golden(gbp, dev):
    _, _, _, G, _ = dev
    s, a, tt = G["pair_states"], G["pair_actions"], G["prop_t"]
    assert_bits_equal(gbp.propagate(0, s, a, tt), G["prop_stance"], what="applyStance")
    assert_bits_equal(gbp.propagate(1, s, None, tt - 0.1), G["prop_flight"], what="applyFlight")
    assert_bits_equal(gbp.propagate(2, s, a, tt), G["prop_stance_rev"], what="applyStanceReverse")


def test_valid_state_action_golden(gbp, dev):
    t, o, T, G, _ = dev
    for phase, key in ((gbp.STANCE, "vs_stance"), (gbp.FLIGHT, "vs_flight")):
        v, fl = t.valid_states(G["vs_states"], phase)
        ok = (fl & gbp.FLAG_OOG) == 0
        assert ok.mean() > 0.9 and (v[ok] == G[key][ok]).all()
        assert ((fl & gbp.FLAG_VALID) == v).all()
    assert (gbp.valid_actions(G["va_actions"]) == G["va_verdict"]).all()


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("adaptive", [0, 1])
def test_validate_pairs_golden(gbp, dev, variant, adaptive):
    t, o, T, G, _ = dev
    if variant == 2 and adaptive:
        with pytest.raises(gbp.GbpError, match="fixed step"):
            t.validate_pairs(G["pair_states"][:4], G["pair_actions"][:4], G["pair_dir"][:4], adaptive=True, variant=2)
        return
    v, fl, sn, tn = t.validate_pairs(G["pair_states"], G["pair_actions"], G["pair_dir"], adaptive=bool(adaptive), variant=variant)
    ing = (fl & gbp.FLAG_OOG) == 0
    assert (ing == G[f"pair_ingrid_{adaptive}"].astype(bool)).all()
    assert (v[ing] == G[f"pair_verdict_{adaptive}"][ing]).all(), "verdict bits differ from the reference"
    assert_bits_equal(sn, G[f"pair_snew_{adaptive}"], where=ing, what="s_new")
    assert_bits_equal(tn, G[f"pair_tnew_{adaptive}"], where=ing, what="t_new")
    # guard band: candidates whose verdict hinges on a margin < 1e-11 m or a probe < 1e-11 m from a grid line are
    # flagged, not silently decided; they must be vanishingly rare (and here still agree with the reference)
    assert (fl & gbp.FLAG_NEAR).mean() < 1e-3
    # work counters equal the oracle's count of the reference's early-exit work
    vo, fo, sno, tno, cnt = o.validate_pairs(G["pair_states"], G["pair_actions"], G["pair_dir"], adaptive=bool(adaptive))
    c = t.validate_counters()
    assert (c["substates"], c["lookups"], c["nanprobes"]) == tuple(int(x) for x in cnt)
    assert c["valid"] == int(vo.sum()) and c["oog"] == int(((fo & po.FLAG_OOG) != 0).sum())
    # including out-of-grid candidates the CUDA path equals the oracle's DEFINED semantics everywhere
    assert (v == vo).all() and (fl & gbp.FLAG_OOG == fo & po.FLAG_OOG).all()
    assert_bits_equal(sn, sno, what="s_new vs oracle")
    assert_bits_equal(tn, tno, what="t_new vs oracle")


def test_sampler_stream_identical(gbp, dev):
    t, o, T, G, name = dev
    n = 5000
    a = gbp.sample_actions(7, 3, 11, n, G["pair_normal"])
    assert np.array_equal(a.view(np.uint64), o.sample_actions(7, 3, 11, n, G["pair_normal"]).view(np.uint64))
    q = t.sample_states(7, (5 << 32) + 4, (1 << 33) + 2, n)
    assert np.array_equal(q.view(np.uint64), o.sample_states(7, (5 << 32) + 4, (1 << 33) + 2, n).view(np.uint64))
    sf, st = G["pair_states"][0], G["pair_states"][1]
    ad = gbp.sample_actions(7, 3, 0, 64, G["pair_normal"], sf, st, 0.5)
    ao = np.stack([o.sample_action_dir(7, 3, i, G["pair_normal"], 0.5, sf, st) for i in range(64)])
    assert np.array_equal(ad.view(np.uint64), ao.view(np.uint64))
    qd = t.sample_states(7, 3, 0, 64, sf, st, 0.5, True)
    qo = np.stack([o.sample_state_dir(7, 3, i, 0.5, True, sf, st) for i in range(64)])
    assert np.array_equal(qd.view(np.uint64), qo.view(np.uint64))


def test_tree_queries_golden(gbp, dev):
    t, o, T, G, _ = dev
    tree = gbp.Tree(2048)
    tree.load(G["nn_verts"])
    assert tree.size() == len(G["nn_verts"])
    idx, dist = tree.nearest(G["nn_q"])
    assert_bits_equal(dist, G["nn_dist"], what="nearest distance")
    u = G["nn_unique"].astype(bool)
    assert (idx[u] == G["nn_idx"][u]).all()
    io, do, _ = o.nearest(G["nn_verts"], G["nn_q"])
    assert (idx == io).all()  # ties included: lowest id, as the oracle defines
    off = 0
    for j, cnt in enumerate(G["near_counts"]):
        ids, total = tree.near(G["nn_q"][j], 3.0)
        assert total == cnt and (ids == G["near_ids"][off:off + cnt]).all()
        off += cnt
    v = G["nn_verts"]
    assert_bits_equal(gbp.distance(0, v[:500], v[500:1000]), G["dist_pose"], what="poseDistance")
    assert_bits_equal(gbp.distance(1, v[:500], v[500:1000]), G["dist_state"], what="stateDistance")
    assert np.allclose(gbp.distance(2, v[:500], v[500:1000]), G["dist_yaw"], rtol=0, atol=1e-12)
    # duplicated vertices: tie -> lowest id
    dup = gbp.Tree(8); dup.load(np.repeat(v[:1], 5, axis=0))
    assert dup.nearest(v[3:4])[0][0] == 0


def test_tree_append_and_gy(gbp, dev):
    _, _, _, G, _ = dev
    verts, parent = G["nn_verts"][:300], G["gy_parent"]
    tree = gbp.Tree(300, verts[0])
    for i in range(1, 300):
        assert tree.append(int(parent[i]), verts[i], np.full(10, float(i))) == i
    with pytest.raises(gbp.GbpError):
        tree.append(0, verts[0], np.zeros(10))  # full
    r = tree.read()
    assert (r["parent"] == parent).all() and (r["states"] == verts).all() and (r["actions"][7] == 7.0).all()
    assert_bits_equal(r["g"], G["gy_g"], what="g values")
    assert np.allclose(r["yaw"], G["gy_y"], rtol=0, atol=1e-10)


def test_attempt_connect_golden(gbp, dev):
    t, o, T, G, _ = dev
    st, sn, an, fl = t.attempt_connect(G["con_existing"], G["con_target"], G["con_dir"])
    ing = (fl & gbp.FLAG_OOG) == 0
    assert (ing == G["con_ingrid"].astype(bool)).all()
    assert (st[ing] == G["con_status"][ing]).all()
    hit = (st != gbp.TRAPPED) & ing
    assert_bits_equal(sn[hit], G["con_snew"][hit], what="attemptConnect s_new")
    assert_bits_equal(an[hit], G["con_anew"][hit], what="attemptConnect a_new")


def test_extend_matches_oracle(gbp, dev):
    """gbp_extend (NN + K candidates + selection + append) against the oracle's newConfig."""
    t, o, T, G, name = dev
    s = G["pair_states"]
    # K x lanes per candidate of k_extend_fused on a 148-SM part: 6 / 256 -> 32, 2000 -> 16, 4096 -> 8, 40000 -> 1; the
    # adaptive step takes the thread-per-candidate kernel
    for K, best, adaptive in ((6, 0, False), (256, 1, False), (256, 0, False), (2000, 0, False), (4096, 1, False), (40000, 1, False),
                              (6, 0, True), (300, 1, True)):
        tree = gbp.Tree(512)
        tree.load(s[:200])
        for trial in range(12 if K <= 256 else 4):
            target = s[300 + trial]
            near, _, _ = o.nearest(s[:200], target[None])
            s_near = s[int(near[0])]
            normal = o.surface_normal(target[:1], target[1:2])[0]
            a = o.sample_actions(9, 4, trial * K, K, normal)
            direction = trial % 2
            vo, _, sno, _, _ = o.validate_pairs(np.repeat(s_near[None], K, 0), a, direction, adaptive=adaptive, nthreads=8)
            d = o.distance(sno, np.repeat(target[None], K, 0), 1)
            base = o.distance(s_near[None], target[None], 1)[0]
            exp = None
            if vo.any():
                j = int(np.argmin(np.where(vo == 1, d, np.inf))) if best else int(np.argmax(vo))
                if d[j] < base:
                    exp = j
            before = tree.size()
            st, nid, chk = tree.extend(t, target, direction, K, best, 9, 4, trial * K, adaptive=adaptive)
            if exp is None:
                assert st == gbp.TRAPPED and tree.size() == before
            else:
                assert st != gbp.TRAPPED and nid == before
                r = tree.read(nid, 1)
                assert_bits_equal(r["states"][0], sno[exp], what="extend s_new")
                assert_bits_equal(r["actions"][0], a[exp], what="extend a_new")
                assert r["parent"][0] == near[0]
                assert st == (gbp.REACHED if d[exp] <= 0.5 else gbp.ADVANCED)
            assert chk == (K if best or not vo.any() else int(np.argmax(vo)) + 1)


def test_connect_matches_oracle(gbp, dev):
    t, o, T, G, _ = dev
    s = G["pair_states"]
    tree = gbp.Tree(256)
    tree.load(s[:100])
    for trial in range(20):
        target = G["con_target"][trial]
        near, _, _ = o.nearest(tree.read()["states"], target[None])
        verts = tree.read()["states"]
        so, sno, ano, flo = o.attempt_connect(verts[int(near[0])][None], target[None], trial % 2)
        before = tree.size()
        st, nid = tree.connect(t, target, trial % 2)
        assert st == so[0]
        if st != gbp.TRAPPED:
            r = tree.read(nid, 1)
            assert nid == before and r["parent"][0] == near[0]
            assert_bits_equal(r["states"][0], sno[0], what="connect s_new")
            assert_bits_equal(r["actions"][0], ano[0], what="connect a_new")


def test_gridmap_ingest(gbp):
    """fast_terrain_map.cpp:31-91 index flip: C-ABI gridmap loader == loadData of the flipped layers."""
    rng = np.random.default_rng(1)
    nx, ny, res = 12, 9, 0.25
    elev = rng.normal(0, 0.1, (nx, ny)).astype(np.float32)
    t = gbp.Terrain.from_gridmap(nx, ny, res, 1.0, -0.5, elev)
    x, y = t.axes()
    assert np.allclose(np.diff(x), res) and x[0] == 1.0 - 0.5 * (nx - 1) * res
    t2 = gbp.Terrain(x, y, elev[::-1, ::-1].astype(np.float64))
    px = rng.uniform(x[0], x[-1] - 1e-9, 500); py = rng.uniform(y[0], y[-1] - 1e-9, 500)
    assert_bits_equal(t.ground_height(px, py)[0], t2.ground_height(px, py)[0], what="gridmap")
    if po.Ref.available():
        r = po.Ref(); r.set_terrain_gridmap(nx, ny, res, 1.0, -0.5, elev)
        rx, ry = r.axes()
        assert (rx == x).all() and (ry == y).all()
        assert_bits_equal(t.ground_height(px, py)[0], r.ground_height(px, py), what="gridmap vs reference")


def test_empty_and_ragged_inputs(gbp, dev):
    t, o, T, G, _ = dev
    v, fl, sn, tn = t.validate_pairs(np.zeros((0, 8)), np.zeros((0, 10)), np.zeros(0, np.uint8))
    assert len(v) == 0
    for n in (1, 31, 33, 65, 127):  # ragged warp tails
        for variant in (1, 2, 3):
            v, fl, sn, tn = t.validate_pairs(G["pair_states"][:n], G["pair_actions"][:n], G["pair_dir"][:n], variant=variant)
            assert (v[(fl & 2) == 0] == G["pair_verdict_0"][:n][(fl & 2) == 0]).all()
    assert len(t.ground_height([], [])[0]) == 0


def test_interp_path_and_curvature_golden(gbp, dev):
    """getInterpPath / calculateMaxCurvature through the C ABI: bit-equal to the reference's vectors and the oracle"""
    t, o, T, G, _ = dev
    s, tt, ph = gbp.interp_path(G["interp_in_states"], G["interp_in_actions"], 0.05)
    assert len(s) == len(G["interp_states"]) and len(ph) == len(s) - 1
    assert_bits_equal(s, G["interp_states"], what="interp states")
    assert_bits_equal(tt, G["interp_t"], what="interp times")
    assert (ph == G["interp_phase"]).all()
    assert gbp.max_curvature(s) == float(G["interp_max_curvature"])
    for dt in (0.03, 0.011, 1.0):
        so, to, po_ = o.interp_path(G["interp_in_states"], G["interp_in_actions"], dt)
        sg, tg, pg = gbp.interp_path(G["interp_in_states"], G["interp_in_actions"], dt)
        assert_bits_equal(sg, so, what=f"interp states dt={dt}")
        assert np.array_equal(tg, to) and np.array_equal(pg, po_)
        assert gbp.max_curvature(sg) == o.max_curvature(so)
    # capacity smaller than the path, empty action list, degenerate plans
    sg, tg, pg = gbp.interp_path(G["interp_in_states"], G["interp_in_actions"], 0.05, cap=7)
    assert len(sg) == 7 and (sg == G["interp_states"][:7]).all()
    s1, t1, p1 = gbp.interp_path(G["interp_in_states"][:1], np.zeros((0, 10)), 0.05)
    assert len(s1) == 1 and (s1[0] == G["interp_in_states"][0]).all() and t1[0] == 0.0 and len(p1) == 0
    assert gbp.max_curvature(G["interp_in_states"][:2]) == 0.0
    with pytest.raises(gbp.GbpError):
        gbp.interp_path(G["interp_in_states"], G["interp_in_actions"], 0.0)


def test_csv_ingest(gbp, dev, tmp_path):
    """The reference's on-disk terrain format (data/<name>/*.csv, rows = y, columns = x): the loadData route equals
    the arrays handed over directly, the GridMap route equals gbp_terrain_create_gridmap on the float-rounded layers."""
    t, o, T, G, name = dev
    X, Y = np.meshgrid(T.x, T.y)  # [ny, nx]
    lay = {"x": X, "y": Y, "z": T.z.T, "dx": T.dx.T, "dy": T.dy.T, "dz": T.dz.T}
    for k, v in lay.items():
        with open(tmp_path / f"{k}data.csv", "w") as f:
            if k == "z":
                f.write("# elevation, rows = y\n")
            for row in v:
                f.write(",".join("nan" if np.isnan(c) else repr(float(c)) for c in row) + "\n")
    tc = gbp.Terrain.from_csv(str(tmp_path))
    assert (tc.nx, tc.ny) == (len(T.x), len(T.y)) and tc.cell_bytes == t.cell_bytes
    xa, ya = tc.axes()
    assert (xa == T.x).all() and (ya == T.y).all()
    px, py = G["probe_x"], G["probe_y"]
    assert_bits_equal(tc.ground_height(px, py)[0], G["probe_h"], what="height via CSV")
    assert_bits_equal(tc.surface_normal(px, py), G["probe_normal"], what="normal via CSV")
    assert (tc.height_is_nan(px, py) == G["probe_nan"]).all()
    # ROS route: uniform square cells only (the reference throws otherwise)
    res = np.float32(T.x[1] - T.x[0])
    if np.float32(T.y[1] - T.y[0]) == res:
        tg = gbp.Terrain.from_csv(str(tmp_path), via_gridmap=True)
        nx, ny = len(T.x), len(T.y)
        xl = T.x[-1] - T.x[0] + res; yl = T.y[-1] - T.y[0] + res
        cx = T.x[0] - 0.5 * res + 0.5 * xl; cy = T.y[0] - 0.5 * res + 0.5 * yl
        flip = lambda a: np.ascontiguousarray(a[::-1, ::-1].astype(np.float32))
        tr = gbp.Terrain.from_gridmap(nx, ny, float(res), float(cx), float(cy), flip(T.z), flip(T.dx), flip(T.dy), flip(T.dz))
        x2, y2 = tg.axes(); x3, y3 = tr.axes()
        assert (x2 == x3).all() and (y2 == y3).all() and np.allclose(x2, T.x, atol=1e-6)
        inside = (px > x2[0]) & (px < x2[-1] - 1e-6) & (py > y2[0]) & (py < y2[-1] - 1e-6)
        h2, h3 = tg.ground_height(px[inside], py[inside])[0], tr.ground_height(px[inside], py[inside])[0]
        assert_bits_equal(h2, h3, what="height via CSV -> GridMap")
    with pytest.raises(gbp.GbpError):
        gbp.Terrain.from_csv(str(tmp_path / "missing"))


def test_nearest_many_queries_tiled(gbp, dev):
    """enough queries to take the 4-queries-per-pass kernel (k_nearest_tiled), with a ragged last tile and exact ties"""
    t, o, T, G, _ = dev
    verts = G["nn_verts"].copy()
    verts[700] = verts[3]; verts[1200] = verts[3]          # duplicated states: the lowest id must win (Appendix B-4)
    rng = np.random.default_rng(1)
    q = np.concatenate([G["vs_states"][:1001], verts[[3, 700, 9]]])   # 1004 queries + exact hits
    q = q[: 1003]                                           # not a multiple of 4
    tree = gbp.Tree(2048)
    tree.load(verts)
    gi, gd = tree.nearest(q)
    oi, od, uniq = o.nearest(verts, q)
    assert_bits_equal(gd, od, what="nearest distance")
    assert (gi[uniq == 1] == oi[uniq == 1]).all()
    d_all = np.sqrt(((verts[None, :, :] - q[:, None, :]) ** 2).sum(-1))
    assert (gi == d_all.argmin(1)).all() or (gd == d_all.min(1)).all()   # ties resolved to the lowest id
    hit = np.nonzero((q == verts[3]).all(1))[0]
    assert len(hit) >= 1 and (gi[hit] == 3).all()


def test_full_tree_is_a_capacity_error_not_trapped(gbp, dev):
    """gbp_extend / gbp_connect on a tree without room for the accepted vertex return GBP_E_CAPACITY and leave the tree
    untouched (TRAPPED means the reference's newConfig / attemptConnect rejected, rrt.cpp:55-66, rrt_connect.cpp:107)"""
    t, o, T, G, name = dev
    s = G["nn_verts"]
    v, _ = o.valid_states(s, po.STANCE)
    s = s[v == 1]
    hit = None
    for i in range(min(len(s) - 1, 40)):  # a (root, target) pair whose extend succeeds on a roomy tree
        roomy = gbp.Tree(4, s[i])
        for j in range(i + 1, min(len(s), i + 20)):
            st, nid, _ = roomy.extend(t, s[j], gbp.FORWARD, 2048, 1, 9, 4, 0)
            if st != gbp.TRAPPED:
                hit = (i, j)
                break
        if hit:
            break
    if hit is None:
        pytest.skip("no successful extend among the sampled pairs on this map")
    full = gbp.Tree(1, s[hit[0]])
    with pytest.raises(gbp.GbpError, match="-3"):
        full.extend(t, s[hit[1]], gbp.FORWARD, 2048, 1, 9, 4, 0)
    assert full.size() == 1
    # connect: a target right next to the root always connects
    near = s[hit[0]].copy(); near[0] += 0.1
    st, _ = gbp.Tree(4, s[hit[0]]).connect(t, near, gbp.FORWARD)
    if st != gbp.TRAPPED:
        with pytest.raises(gbp.GbpError, match="-3"):
            full.connect(t, near, gbp.FORWARD)
        assert full.size() == 1


def test_rotate_grf_and_curvature_exports(gbp, dev):
    """rotate_grf (planning_utils.cpp:198-231) bit-equal to the reference golden G["grf_out"]; calculateCurvature (:884-899)
    bit-equal to calculateMaxCurvature's per-triple value (the oracle's, pinned against the reference's maximum)"""
    t, o, T, G, name = dev
    assert_bits_equal(gbp.rotate_grf(G["grf_n"], G["grf_f"]), G["grf_out"], what="rotate_grf")
    s = G["interp_states"] if "interp_states" in G else G["pair_states"]
    tri = np.concatenate([s[:-2, :2], s[1:-1, :2], s[2:, :2]], axis=1)[:400]
    got = gbp.curvature(tri)
    for i in (0, 7, len(tri) - 1):  # a three-state plan's maximum curvature IS the curvature of its only triple
        three = np.zeros((3, 8)); three[:, :2] = tri[i].reshape(3, 2)
        want = o.max_curvature(three)
        assert got[i] == want or (np.isnan(got[i]) and want == 0.0)  # std::max never takes a NaN (collinear points)
    deg = gbp.curvature(np.array([[0, 0, 1, 0, 2, 0.0], [1, 1, 1, 2, 1, 3.0], [0, 0, 0, 0, 0, 0.0]]))
    assert (deg == 0).all()
