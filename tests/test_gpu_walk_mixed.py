"""Parity of the headline kernel pair (k_walk_mixed + k_validate_redo + k_pair_outputs) — the path fp32 maps on
uniform axes without NaN take, i.e. bench.py's workload — through the C ABI, against the oracle restatement and the
golden vectors minted from the unmodified reference (tests/golden/golden_synth_mixed.npz).

What is specific to this path and therefore tested here: the texture-gather fetch (and its 4-load twin), the
polynomial-segment cursor (take-off / reverse-stance segment changes, pure-stance and zero-flight primitives, long
connect primitives), warp-level refill over the 16-candidate TMA ring (ragged tails, batches smaller than a warp,
many trips per warp), the border zone and near-threshold candidates that must be handed to the fp64 redo pass, and the
adaptive-step instantiation.  Everything is compared bit for bit: verdicts, flags, s_new, t_new and the k / L / NaN
work counters of the reference's early-exit semantics."""
import os

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
def env(gbp):
    T = load_terrain("synth_mixed")
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    f = t.flags()
    assert f["uniform_axes"] and f["mixed_precision"] and f["texture_gather"], f
    return t, po.Oracle(T), T


def candidates(o, n, seed, gentle=0.3):
    """n (state, action, direction) triples: valid STANCE start states anywhere on the map (border zone included),
    sampled actions, a share of gentle ones so that the fully-valid branches (take-off, landing, exact start) are hit."""
    states = []
    idx0 = 0
    while sum(len(x) for x in states) < n:
        q = o.sample_states(seed, 21, idx0, 4 * n + 1024)
        idx0 += len(q)
        v, _ = o.valid_states(q, po.STANCE)
        states.append(q[v == 1])
    s = np.concatenate(states)[:n]
    a = o.sample_actions(seed, 22, 0, n, (0.02, -0.01, 0.995))
    rng = np.random.default_rng(seed)
    g = rng.random(n) < gentle
    a[g, :6] = rng.normal(0, 1.0, (int(g.sum()), 6))
    a[g, 8:] = rng.normal(0, 0.5, (int(g.sum()), 2))
    a[g, 7] = rng.uniform(0, 0.2, int(g.sum()))
    d = rng.integers(0, 2, n).astype(np.uint8)
    return s, a, d


def check(gbp, t, o, s, a, d, adaptive=False, variant=3):
    v, fl, sn, tn = t.validate_pairs(s, a, d, adaptive=adaptive, variant=variant)
    c = t.validate_counters()
    vo, fo, sno, tno, cnt = o.validate_pairs(s, a, d, adaptive=adaptive)
    assert (v == vo).all(), f"{int((v != vo).sum())} verdicts differ"
    assert ((fl & gbp.FLAG_OOG) == (fo & po.FLAG_OOG)).all() and ((fl & gbp.FLAG_VALID) == v).all()
    assert_bits_equal(sn, sno, what="s_new")
    assert_bits_equal(tn, tno, what="t_new")
    assert (c["substates"], c["lookups"], c["nanprobes"]) == tuple(int(x) for x in cnt), (c, cnt)
    assert c["valid"] == int(vo.sum())
    return v


@pytest.mark.parametrize("adaptive", [False, True])
def test_large_batch_matches_oracle(gbp, env, adaptive):
    t, o, T = env
    s, a, d = candidates(o, 60000, seed=31)
    v = check(gbp, t, o, s, a, d, adaptive=adaptive)
    assert 0.01 < v.mean() < 0.9  # both outcomes well represented


@pytest.mark.parametrize("n", [1, 5, 15, 16, 17, 31, 33, 63, 64, 65, 127, 2049])
def test_ragged_sizes(gbp, env, n):
    t, o, T = env
    s, a, d = candidates(o, 2049, seed=32)
    check(gbp, t, o, s[:n], a[:n], d[:n])


def test_segment_edge_cases(gbp, env):
    """zero / tiny flight, stance times that hit or miss the sample grid, long connect-like primitives, both directions"""
    t, o, T = env
    s, a, d = candidates(o, 4096, seed=33, gentle=1.0)
    n = len(s)
    a[0:512, 7] = 0.0                                      # pure stance (CONNECT_STANCE primitives)
    a[512:1024, 7] = 1e-9                                  # one flight sample
    a[1024:1536, 7] = 0.05                                 # t < t_f boundary of the flight loop
    a[1536:2048, 6] = np.linspace(0.051, 0.9, 512)         # stance times off the 0.05 grid
    a[2048:2560, 6] = 0.05 * np.arange(1, 513).clip(1, 40)  # on the grid, up to 2 s
    a[2560:3072, :6] *= 0.05; a[2560:3072, 6] = 3.0         # long gentle primitives (60 stance samples)
    a[3072:3200, 6] = 0.0                                  # t_s = 0: 1 / (6 t_s) is infinite
    a[3200:3328, 6] = -0.1                                 # negative stance time
    a[3328:3456, 7] = -0.1                                 # negative flight time
    for adaptive in (False, True):
        check(gbp, t, o, s, a, d, adaptive=adaptive)
    assert n == 4096


def test_variants_agree(gbp, env):
    t, o, T = env
    s, a, d = candidates(o, 6000, seed=34)
    base = t.validate_pairs(s, a, d, variant=3)
    for variant in (1, 2):
        other = t.validate_pairs(s, a, d, variant=variant)
        assert (base[0] == other[0]).all() and (base[1] == other[1]).all()
        assert_bits_equal(base[2], other[2], what=f"s_new variant {variant}")
        assert_bits_equal(base[3], other[3], what=f"t_new variant {variant}")


def test_four_load_fetch_path_is_identical(gbp, env):
    """GBP_NO_TEX=1 at creation keeps the LDG form of the mixed evaluator: same bits everywhere"""
    t, o, T = env
    os.environ["GBP_NO_TEX"] = "1"
    try:
        t2 = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    finally:
        del os.environ["GBP_NO_TEX"]
    f = t2.flags()
    assert f["mixed_precision"] and not f["texture_gather"]
    s, a, d = candidates(o, 20000, seed=35)
    for adaptive in (False, True):
        check(gbp, t2, o, s, a, d, adaptive=adaptive)


def test_golden_from_reference(gbp, env):
    """the committed vectors of the unmodified reference on this terrain, through the default (texture) walk"""
    t, o, T = env
    G = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_synth_mixed.npz")))
    for adaptive in (0, 1):
        v, fl, sn, tn = t.validate_pairs(G["pair_states"], G["pair_actions"], G["pair_dir"], adaptive=bool(adaptive))
        ing = (fl & gbp.FLAG_OOG) == 0
        assert (ing == G[f"pair_ingrid_{adaptive}"].astype(bool)).all()
        assert (v[ing] == G[f"pair_verdict_{adaptive}"][ing]).all()
        assert_bits_equal(sn, G[f"pair_snew_{adaptive}"], where=ing, what="s_new")
        assert_bits_equal(tn, G[f"pair_tnew_{adaptive}"], where=ing, what="t_new")


def test_device_pointer_call_and_alignment(gbp, env):
    """gbp_validate_pairs_dev on torch device buffers; misaligned pointers are rejected, not mis-copied"""
    import torch
    t, o, T = env
    s, a, d = candidates(o, 5000, seed=36)
    dev = torch.device("cuda:0")
    ds, da, dd = torch.from_numpy(s).to(dev), torch.from_numpy(a).to(dev), torch.from_numpy(d).to(dev)
    n = len(s)
    dv = torch.zeros(n, dtype=torch.uint8, device=dev); df = torch.zeros_like(dv)
    dsn = torch.zeros((n, 8), dtype=torch.float64, device=dev); dtn = torch.zeros(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    t.validate_pairs_dev(n, ds.data_ptr(), da.data_ptr(), dd.data_ptr(), 0, 0, dv.data_ptr(), df.data_ptr(), dsn.data_ptr(), dtn.data_ptr(), 0)
    torch.cuda.synchronize()
    vo, fo, sno, tno, _ = o.validate_pairs(s, a, d)
    assert (dv.cpu().numpy() == vo).all()
    assert_bits_equal(dsn.cpu().numpy(), sno, what="s_new (device call)")
    assert_bits_equal(dtn.cpu().numpy(), tno, what="t_new (device call)")
    with pytest.raises(gbp.GbpError, match="aligned"):
        t.validate_pairs_dev(n - 1, ds.data_ptr() + 8, da.data_ptr(), dd.data_ptr(), 0, 0, dv.data_ptr(), df.data_ptr(), dsn.data_ptr(), dtn.data_ptr(), 0)


@pytest.mark.parametrize("n", [1, 17, 1000, 33333])
def test_no_write_outside_the_outputs(gbp, env, n):
    """guard bands around every output buffer of the device-pointer call stay untouched (compute-sanitizer is not
    available on the GPU pool: this is the bounds check the kernels get)"""
    import torch
    t, o, T = env
    s, a, d = candidates(o, n, seed=37)
    dev = torch.device("cuda:0")
    pad = 4096  # bytes, keeps the 16-byte alignment the TMA copies need
    def guarded(nbytes):
        buf = torch.full((pad + nbytes + pad,), 0xA5, dtype=torch.uint8, device=dev)
        return buf, buf.data_ptr() + pad
    ds, da, dd = torch.from_numpy(s).to(dev), torch.from_numpy(a).to(dev), torch.from_numpy(d).to(dev)
    bufs = {k: guarded(sz) for k, sz in (("v", n), ("f", n), ("sn", 64 * n), ("tn", 8 * n))}
    for adaptive in (0, 1):
        t.validate_pairs_dev(n, ds.data_ptr(), da.data_ptr(), dd.data_ptr(), adaptive, 0, bufs["v"][1], bufs["f"][1], bufs["sn"][1], bufs["tn"][1], 0)
        torch.cuda.synchronize()
        for k, (buf, _) in bufs.items():
            h = buf.cpu().numpy()
            assert (h[:pad] == 0xA5).all() and (h[-pad:] == 0xA5).all(), f"{k}: guard band overwritten"
        vo, fo, sno, tno, _ = o.validate_pairs(s, a, d, adaptive=bool(adaptive))
        assert (bufs["v"][0].cpu().numpy()[pad:pad + n] == vo).all()
        sn = bufs["sn"][0].cpu().numpy()[pad:pad + 64 * n].view(np.float64).reshape(n, 8)
        assert_bits_equal(sn, sno, what="s_new")


def test_fp64_map_through_rounded_texture_copy(gbp):
    """fp64 heights that are not fp32 numbers (the shipped CSV maps' class), lifted to 3.9 m where the fp32 rounding of
    the texture copy is largest (2.4e-7 m): the mixed walk applies and every result still equals the oracle's; above
    4 m the rounded copy is not used at all."""
    T0 = load_terrain("synth_mixed")
    rng = np.random.default_rng(3)
    z = T0.z + 3.83 + rng.uniform(0, 1e-6, T0.z.shape)       # no longer fp32-representable
    assert (z.astype(np.float32).astype(np.float64) != z).mean() > 0.9 and np.abs(z).max() < 4.0
    T = po.Terrain(T0.x, T0.y, z, T0.dx, T0.dy, T0.dz)
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    assert t.cell_bytes == 8 and t.flags()["mixed_precision"] and t.flags()["texture_gather"]
    o = po.Oracle(T)
    s, a, d = candidates(o, 30000, seed=41)
    for adaptive in (False, True):
        check(gbp, t, o, s, a, d, adaptive=adaptive)
    # near-threshold poses: body heights swept in 1e-7 m steps across the belly-clearance and leg-reach limits
    q = s[:2000].copy()
    h = o.ground_height(q[:, 0], q[:, 1])[0]
    q[:, 3:] = 0; q[:, 3] = 0.3
    q[:1000, 2] = h[:1000] + 0.125 + np.arange(-500, 500) * 1e-7
    q[1000:, 2] = h[1000:] + 0.45 + np.arange(-500, 500) * 1e-7
    for phase in (gbp.STANCE, gbp.FLIGHT):
        vg, fg = t.valid_states(q, phase)
        vo, fo = o.valid_states(q, phase)
        assert (vg == vo).all()
    t2 = gbp.Terrain(T.x, T.y, T.z + 1.0)
    assert not t2.flags()["mixed_precision"] and not t2.flags()["texture_gather"]


def test_large_ragged_batch_device_call(gbp, env):
    """150 001 candidates (ragged for every tile size in use) through the device-pointer call: guard bands around s_new,
    results bit-equal to the oracle"""
    import torch
    t, o, T = env
    n = 150001
    s, a, d = candidates(o, n, seed=43)
    dev = torch.device("cuda:0")
    ds, da, dd = torch.from_numpy(s).to(dev), torch.from_numpy(a).to(dev), torch.from_numpy(d).to(dev)
    pad = 4096
    buf = torch.full((pad + 64 * n + pad,), 0xA5, dtype=torch.uint8, device=dev)
    dv = torch.zeros(n, dtype=torch.uint8, device=dev); df = torch.zeros_like(dv); dtn = torch.zeros(n, dtype=torch.float64, device=dev)
    t.validate_pairs_dev(n, ds.data_ptr(), da.data_ptr(), dd.data_ptr(), 0, 0, dv.data_ptr(), df.data_ptr(), buf.data_ptr() + pad, dtn.data_ptr(), 0)
    torch.cuda.synchronize()
    h = buf.cpu().numpy()
    assert (h[:pad] == 0xA5).all() and (h[-pad:] == 0xA5).all(), "guard band overwritten"
    vo, fo, sno, tno, _ = o.validate_pairs(s, a, d, nthreads=8)
    assert (dv.cpu().numpy() == vo).all()
    assert_bits_equal(h[pad:pad + 64 * n].view(np.float64).reshape(n, 8), sno, what="s_new vs oracle")
    assert_bits_equal(dtn.cpu().numpy(), tno, what="t_new vs oracle")


def test_hard_thresholds_are_guarded_and_cliffs_disable_the_mixed_walk(gbp, env):
    """isValidState's speed (> V_MAX) and pitch (>= P_MAX) tests are hard comparisons on poses that, inside a pair check, come
    from FMA forms of the primitives: a value within 1e-12 of its threshold is flagged GBP_FLAG_NEAR (and never decided by the
    mixed-precision evaluator); verdicts on exact inputs still equal the oracle's.  A map with cliffs above 4 m per cell is
    outside the mixed evaluator's error budget and keeps the fp64 walk."""
    t, o, T = env
    q = candidates(o, 64, seed=77)[0]
    q[:, 3:6] = 0.0
    q[:32, 3] = 2.0 + np.linspace(-4e-13, 4e-13, 32)             # speed on both sides of V_MAX = 2
    q[32:, 3] = 0.5; q[32:, 6] = 1.0 + np.linspace(-4e-13, 4e-13, 32)  # pitch on both sides of P_MAX = 1
    for phase in (gbp.STANCE, gbp.FLIGHT):
        vg, fg = t.valid_states(q, phase)
        vo, fo = o.valid_states(q, phase)
        assert (vg == vo).all()
        assert ((fg & gbp.FLAG_NEAR) != 0).sum() >= 60
    far = q.copy(); far[:, 3] = 1.0; far[:, 6] = 0.2
    assert not (t.valid_states(far, gbp.STANCE)[1] & gbp.FLAG_NEAR).any()
    z = T.z.copy()
    z[T.nx // 2:, :] += 10.0                                          # a 10 m cliff across the map
    t2 = gbp.Terrain(T.x, T.y, z, T.dx, T.dy, T.dz)
    assert not t2.flags()["mixed_precision"]
    s, a, d = candidates(po.Oracle(po.Terrain(T.x, T.y, z, T.dx, T.dy, T.dz)), 3000, seed=78)
    vg = t2.validate_pairs(s, a, d)[0]
    vo = po.Oracle(po.Terrain(T.x, T.y, z, T.dx, T.dy, T.dz)).validate_pairs(s, a, d)[0]
    assert (vg == vo).all()
