"""The publisher node's procedural terrains: TerrainMapPublisher::createOwnMap / createMap
(terrain_map_publisher.cpp:34-231, :253-286).  CPU tests pin the oracle restatement against an independent numpy
restatement of the deterministic parts (axes, findXYIndex, fill order, grid_map flip) and of the Philox draw spec; the
GPU tests compare the device generator with the oracle bit for bit and feed the result through the GridMap ingest."""
import math

import numpy as np
import pytest

import pyoracle as po
from conftest import assert_bits_equal

REF_RECTS = np.array(  # terrain_map_publisher.cpp:107-127
    [[-np.finfo(float).max, -np.finfo(float).max, np.finfo(float).max, np.finfo(float).max, 0, 0.01],
     [8.13, -4, 8.42, 4, 0.158, 0.01], [8.42, -4, 8.71, 4, 0.316, 0.01], [8.71, -4, 10.5, 4, 0.474, 0.01],
     [0.75, -3.15, 2.05, -2.35, 0.6, 0.1], [4.25, -2.4, 5.4, -1.75, 0.5, 0.05], [2.9, -0.6, 3.25, 1.15, 0.158, 0.01],
     [4.9, -0.5, 5.3, 0.35, -0.3, 0.01], [6.5, 0.45, 7.2, 1.05, 0.7, 0.07], [0.65, 2.95, 1.15, 3.75, 0.3, 0.08],
     [4.4, 2.8, 5.7, 3.55, 0.65, 0.04], [7.5, -2.6, 9.45, -1.15, 0.68, 0.06], [6.2, 1.1, 9.2, 2.3, -0.2, 0.06]])
EDGE_RECTS = np.array(  # constant fills, NaN hole, degenerate / outside rectangles, rectangles ending on the last node
    [[0.123, 0.456, 0.789, 0.9, 0.05, 0.5], [0.0, 0.0, 0.3, 0.3, 0.2, 0.0], [0.5, 0.5, 0.75, 0.6, np.nan, 0.0], [2.0, 0.0, 1.0, 1.0, 9.0, 0.1],
     [0.3, 0.3, 0.3, 0.9, 9.0, 0.1], [50.0, 0.0, 60.0, 1.0, 9.0, 0.1], [-9.0, -9.0, -8.0, 9.0, 9.0, 0.1],
     [1.45, -0.2, 1.45, 0.0, 9.0, 0.1], [1.2, 0.9, 1.45, 0.95, -0.1, 0.03], [-1.0, -1.0, -0.5, -0.2, 0.4, 0.05],
     [-0.5, -0.2, 0.0, 0.0, 0.3, 0.02]])
EDGE_GEOM = dict(x_size=40, y_size=24, x_start=-0.5, y_start=-0.2, res=0.05)


def np_axes(n, start, res):
    ax, v = np.zeros(n), start
    rnd = lambda t: math.floor(abs(t) + 0.5) * (1 if t >= 0 else -1)  # C round(): halves away from zero
    for i in range(n):
        ax[i] = rnd(v * 100) / 100
        v = rnd((v + res) * 100) / 100
    return ax


def np_range(xa, ya, r):
    """findXYIndex (terrain_map_publisher.cpp:178-231) after the early returns of :153-156"""
    x1, y1, x2, y2 = r[:4]
    if x1 > xa[-1] or x2 < xa[0] or y1 > ya[-1] or y2 < ya[0] or x1 >= x2 or y1 >= y2:
        return 0, 0, 0, 0
    lo = lambda ax, v: 0 if v <= ax[0] else min(int(np.searchsorted(ax, v, side="right")) - 1, len(ax) - 1)
    hi = lambda ax, v: len(ax) if v >= ax[-1] else int(np.searchsorted(ax, v, side="right"))
    return lo(xa, x1), lo(ya, y1), hi(xa, x2), hi(ya, y2)


def py_draw(o, seed, rect_no, cell, mu, delta):
    """the TERRAIN cell of the Philox spec (include/gbp_b200.h) from the oracle's primitive stream"""
    if not (delta > 0) or mu != mu:
        return mu
    for b in range(16):
        ua, ub = o.uniforms(seed, rect_no, cell, 3, 2 * b, 2)
        r = math.sqrt(-2.0 * o.det_log(1.0 - ua)); sn, cs = o.det_sincos(6.283185307179586 * ub)
        for z in (r * cs, r * sn):
            val = z * delta + mu
            if not (val < mu - delta or val > mu + delta):
                return val
    return mu


def np_own_map(o, seed, x_size, y_size, x_start, y_start, res, rects):
    xa, ya = np_axes(x_size, x_start, res), np_axes(y_size, y_start, res)
    z = np.zeros((y_size, x_size))
    for k, r in enumerate(rects):
        a, b, c, d = np_range(xa, ya, r)
        for i in range(b, d):
            for j in range(a, c):
                z[i, j] = py_draw(o, seed, k, i * x_size + j, r[4], r[5])
    return xa, ya, z


def test_own_map_axes_and_ranges():
    xa = po.own_map_axes(221, -0.5, 0.05); ya = po.own_map_axes(161, -4.0, 0.05)
    assert np.array_equal(xa, np_axes(221, -0.5, 0.05)) and np.array_equal(ya, np_axes(161, -4.0, 0.05))
    assert xa[0] == -0.5 and xa[-1] == 10.5 and ya[0] == -4.0 and ya[-1] == 4.0
    assert np.array_equal(xa, np.round(np.arange(221) * 5 - 50) / 100)  # every node is the nearest double of a centimetre value
    for r in REF_RECTS:
        assert tuple(po.own_map_range(xa, ya, r)) == np_range(xa, ya, r)
    assert tuple(po.own_map_range(xa, ya, REF_RECTS[0])) == (0, 0, 221, 161)
    assert tuple(po.own_map_range(xa, ya, REF_RECTS[1])) == (172, 0, 179, 161)  # 8.13 in [8.10, 8.15), 8.42 in [8.40, 8.45)
    xe, ye = np_axes(40, -0.5, 0.05), np_axes(24, -0.2, 0.05)
    for r in EDGE_RECTS:
        assert tuple(po.own_map_range(xe, ye, r)) == np_range(xe, ye, r)


def test_own_map_matches_numpy_restatement():
    o = po.Oracle()
    elev, geom = po.own_map(7, rects=EDGE_RECTS, **EDGE_GEOM)
    xa, ya, z = np_own_map(o, 7, rects=EDGE_RECTS, **EDGE_GEOM)
    want = z.astype(np.float32)[::-1, ::-1].T  # grid_map index (i, j) = z_data[ny-1-j][nx-1-i] (:88-93)
    assert np.array_equal(elev, want, equal_nan=True)
    assert np.isnan(elev).sum() > 0 and (elev == np.float32(0.2)).sum() > 0
    assert geom[0] == 0.05 and geom[1] == xa[0] - 0.5 * 0.05 + 0.5 * ((-0.5 + 0.05 * 39) - -0.5 + 0.05)
    # the reference's own configuration: spot-check 300 cells of the 221 x 161 map, and the truncation bounds everywhere
    elev, geom = po.own_map(1)
    assert elev.shape == (221, 161) and tuple(geom) == (0.05, 5.0, 0.0)
    xa, ya = np_axes(221, -0.5, 0.05), np_axes(161, -4.0, 0.05)
    z = elev[::-1, ::-1].T.astype(np.float64)
    last = np.zeros((161, 221), int)
    for k, r in enumerate(REF_RECTS):
        a, b, c, d = np_range(xa, ya, r)
        last[b:d, a:c] = k
    mu, dl = REF_RECTS[last, 4], REF_RECTS[last, 5]
    assert (np.abs(z - mu) <= dl + 1e-7).all() and len(np.unique(last)) == 13
    rng = np.random.default_rng(0)
    for i, j in zip(rng.integers(0, 161, 300), rng.integers(0, 221, 300)):
        k = last[i, j]
        assert elev[220 - j, 160 - i] == np.float32(py_draw(o, 1, int(k), int(i * 221 + j), REF_RECTS[k, 4], REF_RECTS[k, 5]))
    assert not np.array_equal(po.own_map(2)[0], elev)


def test_default_map_restatement():
    elev, geom = po.default_map()
    assert tuple(geom) == (0.2, 4.0, 0.0) and elev.shape == (60, 25)
    i, j = np.meshgrid(np.arange(60), np.arange(25), indexing="ij")
    px, py = 4.0 + (0.5 * 59 - i) * 0.2, 0.0 + (0.5 * 24 - j) * 0.2
    assert np.array_equal(elev, np.where((px - 2) ** 2 + py ** 2 <= 0.25, np.float32(0.1), np.float32(0)))
    assert (elev > 0).sum() == 20


@pytest.fixture(scope="module")
def gbp():
    import __graft_entry__ as entry
    entry.build()
    import global_body_planner_b200 as g
    if g.device_count() == 0:
        pytest.skip("no CUDA device")
    return g


@pytest.mark.gpu
def test_own_map_device_generator_bit_equal(gbp):
    for seed, kw in ((1, {}), (99, {}), (7, dict(rects=EDGE_RECTS, **EDGE_GEOM)), (3, dict(rects=REF_RECTS[:1], x_size=640, y_size=512))):
        eg, gg = gbp.own_map_layer(seed, **kw)
        eo, go = po.own_map(seed, **kw)
        assert np.array_equal(eg.view(np.uint32), eo.view(np.uint32)), f"seed {seed}: elevation layer differs from the oracle"
        assert np.array_equal(gg, go)
    with pytest.raises(gbp.GbpError):
        gbp.own_map_layer(1, x_size=1)


@pytest.mark.gpu
def test_own_map_terrain_through_gridmap_ingest(gbp):
    """createOwnMap -> GridMap -> loadDataFromGridMap: the terrain equals the GridMap ingest of the oracle's layer, and
    (when the reference is built) the unmodified FastTerrainMap::loadDataFromGridMap on it."""
    for seed, kw in ((1, {}), (7, dict(rects=EDGE_RECTS, **EDGE_GEOM))):
        t = gbp.Terrain.own_map(seed, **kw)
        eo, go = po.own_map(seed, **kw)
        nx, ny = eo.shape
        t2 = gbp.Terrain.from_gridmap(nx, ny, go[0], go[1], go[2], eo)
        x, y = t.axes(); x2, y2 = t2.axes()
        assert np.array_equal(x, x2) and np.array_equal(y, y2)
        assert np.allclose(x, np_axes(nx, kw.get("x_start", -0.5), 0.05), atol=1e-12)
        rng = np.random.default_rng(seed)
        px = rng.uniform(x[0], x[-1] - 1e-9, 4000); py = rng.uniform(y[0], y[-1] - 1e-9, 4000)
        h, h2 = t.ground_height(px, py)[0], t2.ground_height(px, py)[0]
        assert np.array_equal(h.view(np.uint64), h2.view(np.uint64))
        assert np.array_equal(t.height_is_nan(px, py), t2.height_is_nan(px, py))
        nrm = t.surface_normal(px, py)  # no dx / dy / dz layers: bilinear interpolation of (0, 0, 1), not renormalised
        assert_bits_equal(nrm, t2.surface_normal(px, py), what="normals of an own map")
        assert (nrm[:, :2] == 0).all() and np.allclose(nrm[:, 2], 1.0, atol=1e-12)
        if po.Ref.available():
            r = po.Ref(); r.set_terrain_gridmap(nx, ny, go[0], go[1], go[2], eo)
            rx, ry = r.axes()
            assert np.array_equal(rx, x) and np.array_equal(ry, y)
            ok = ~np.isnan(h)
            assert_bits_equal(h[ok], r.ground_height(px, py)[ok], what="own map vs reference loadDataFromGridMap")
    # the default 221 x 161 box world is plannable terrain: mixed-precision class, states validate like the oracle's
    t = gbp.Terrain.own_map(1)
    eo, go = po.own_map(1)
    x, y = t.axes()
    T = po.Terrain(x, y, eo[::-1, ::-1].astype(np.float64))
    o = po.Oracle(T)
    q = t.sample_states(5, 1, 0, 20000)
    v, _ = t.valid_states(q, gbp.STANCE)
    assert np.array_equal(v, o.valid_states(q, po.STANCE)[0]) and 0 < v.sum() < len(v)


@pytest.mark.gpu
def test_default_map(gbp):
    t = gbp.Terrain.default_map()
    eo, go = po.default_map()
    t2 = gbp.Terrain.from_gridmap(60, 25, go[0], go[1], go[2], eo, np.zeros_like(eo), np.zeros_like(eo), np.ones_like(eo))
    x, y = t.axes(); x2, y2 = t2.axes()
    assert np.array_equal(x, x2) and np.array_equal(y, y2) and (t.nx, t.ny) == (60, 25)
    rng = np.random.default_rng(0)
    px = rng.uniform(x[0], x[-1] - 1e-9, 2000); py = rng.uniform(y[0], y[-1] - 1e-9, 2000)
    assert np.array_equal(t.ground_height(px, py)[0], t2.ground_height(px, py)[0])
    assert t.ground_height([2.0], [0.0])[0][0] > 0.09 and t.ground_height([6.0], [1.0])[0][0] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["default", "create:1"])
def test_driver_plans_on_generated_terrain(gbp, tmp_path, source):
    """gbp_plan on the publisher's procedural sources (the `map_data_source` switch of terrain_map_publisher.cpp:417-428):
    (0,0) -> (8,0); every primitive of the returned plan is re-validated by the oracle on the oracle's copy of the map."""
    import os
    import subprocess
    from conftest import ROOT
    if source == "default":
        t = gbp.Terrain.default_map(); eo, _ = po.default_map()
    else:
        t = gbp.Terrain.own_map(1); eo, _ = po.own_map(1)
    x, y = t.axes()
    o = po.Oracle(po.Terrain(x, y, eo[::-1, ::-1].astype(np.float64)))
    exe = os.path.join(ROOT, "global_body_planner_b200", "gbp_plan")
    disc = str(tmp_path / "disc.csv")
    r = subprocess.run([exe, source, "--height", "0.30", "--num-calls", "1", "--replan-time-limit", "0.3", "--max-time-solve", "60",
                        "--discrete-out", disc, "--quiet"], capture_output=True, text=True, timeout=170)
    assert r.returncode == 0, r.stderr
    rows = np.loadtxt(disc, delimiter=",", ndmin=2)
    assert len(rows) >= 2, f"no plan on the {source} terrain within 60 s"
    ss, aa = rows[:, :8], rows[:-1, 8:]
    assert np.allclose(ss[0, :2], [0, 0]) and np.allclose(ss[-1, :2], [8, 0])
    end = o.apply_flight(o.apply_stance(ss[:-1], aa, aa[:, 6]), aa[:, 7])
    assert np.abs(end - ss[1:]).max() < 1e-9
    vf = o.validate_pairs(ss[:-1], aa, np.zeros(len(aa), np.uint8))[0]
    vr = o.validate_pairs(ss[1:], aa, np.ones(len(aa), np.uint8))[0]
    assert ((vf == 1) | (vr == 1)).all(), "a primitive of the returned plan is invalid under the oracle in both directions"
