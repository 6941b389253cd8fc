"""Parity of the narrow-wire sample + validate call (gbp_sample_validate[_dev]: k_walk_sv + k_sv_fp64 + the compaction
kernels) — newConfig's unit of work, rrt.cpp:34-50 — through the C ABI.

compact results == full results == oracle: the verdict bit of candidate i must equal the verdict the dense call
(gbp_validate_pairs) and the oracle give for (table[state_idx[i]], ACTION cell idx0 + i, direction[i]); the rows of the
valid candidates (index ascending, s_new, t_new, action) must equal the dense call's s_new / t_new and the sampled action
bit for bit; the k / L / NaN-probe work counters must equal the oracle's.  Covered: every map class (mixed-precision walk
with texture gather and 4-load fetch, fp64 redo, the general fp64 kernel on NaN / non-uniform / shipped maps), explicit
row numbers with repeats (K candidates per vertex, as in newConfig) and implicit ones with a row offset, per-candidate and
constant directions, adaptive step, directional action sampling, tilted surface normals, ragged sizes around the 32-candidate
production batch and the verdict-word boundary, empty calls, a capped valid list, the device-pointer entry."""
import ctypes as C

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


@pytest.fixture(scope="module", params=["synth_mixed", "rough_terrain", "slope", "synth_nan"])
def env(gbp, request):
    T = load_terrain(request.param)
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    return request.param, t, po.Oracle(T), T


@pytest.fixture(scope="module", params=["synth_mixed", "rough_terrain"])
def env_sizes(gbp, request):
    """size handling does not depend on the map: one mixed-precision map and one shipped fp64 map"""
    T = load_terrain(request.param)
    t = gbp.Terrain(T.x, T.y, T.z, T.dx, T.dy, T.dz)
    return request.param, t, po.Oracle(T), T


def vertex_table(o, rows, seed):
    """valid STANCE states anywhere on the map (the tree vertices candidates start from)"""
    out, idx0 = [], 0
    while sum(len(x) for x in out) < rows:
        q = o.sample_states(seed, 41, idx0, 4 * rows + 1024)
        idx0 += len(q)
        v, _ = o.valid_states(q, po.STANCE)
        out.append(q[v == 1])
    return np.concatenate(out)[:rows]


def check(gbp, t, o, table, n, seed=5, stream=9, idx0=0, state_idx=None, direction=None, direction0=0, row0=0, adaptive=False,
          normal=(0.0, 0.0, 1.0), target=None, thresh=0.0, valid_cap=None, states_valid=False, packed=False):
    tab = gbp.States(table)
    p = gbp.sv_params(seed, stream, idx0, normal, adaptive, direction0, target, thresh, row0, states_valid, direction_in_row=packed)
    if packed:  # one 4-byte word per candidate: row | direction << 31
        r = t.sample_validate(tab, n, p, gbp.pack_rows(state_idx, direction), None, valid_cap=valid_cap)
    else:
        r = t.sample_validate(tab, n, p, state_idx, direction, valid_cap=valid_cap)
    # the same candidates, spelled out
    rows = np.asarray(state_idx, dtype=np.int64) if state_idx is not None else row0 + np.arange(n)
    s = table[rows]
    d = np.asarray(direction, dtype=np.uint8) if direction is not None else np.full(n, direction0, np.uint8)
    if target is None:
        a = o.sample_actions(seed, stream, idx0, n, normal)
    else:  # getRandomAction(surf_norm, direction, flag, thresh, s, s_near): FORWARD from s_near to s, REVERSE from s to s_near
        a = np.zeros((n, 10))
        for i in range(n):
            sf, st = (s[i], target) if d[i] == 0 else (target, s[i])
            a[i] = o.sample_action_dir(seed, stream, idx0 + i, normal, thresh, sf, st)
    vo, fo, sno, tno, cnt = o.validate_pairs(s, a, d, adaptive=adaptive)
    assert np.array_equal(r["verdict"], vo), f"{int((r['verdict'] != vo).sum())} of {n} verdict bits differ from the oracle"
    assert r["n_valid"] == int(vo.sum())
    c = r["counters"]
    assert (c["substates"], c["lookups"], c["nanprobes"]) == tuple(int(x) for x in cnt), (c, cnt)
    assert c["oog"] == int(((fo & po.FLAG_OOG) != 0).sum())
    assert ((r["flags"] & gbp.FLAG_OOG) == (fo & po.FLAG_OOG)).all() and ((r["flags"] & gbp.FLAG_VALID) == vo).all()
    want = np.flatnonzero(vo)
    m = len(want) if valid_cap is None else min(len(want), valid_cap)
    assert np.array_equal(r["index"], want[:m]), "valid rows are not the ascending list of valid candidates"
    assert_bits_equal(r["s_new"], sno[want[:m]], what="valid s_new")
    assert_bits_equal(r["t_new"], tno[want[:m]], what="valid t_new")
    assert_bits_equal(r["action"], a[want[:m]], what="valid action")
    if n:  # and the dense call of the product agrees too (compact == full)
        vd, fd, snd, tnd = t.validate_pairs(s, a, d, adaptive=adaptive)
        assert np.array_equal(vd, r["verdict"])
        assert_bits_equal(r["s_new"], snd[want[:m]], what="valid s_new vs dense call")
    # spare bits of the last verdict word stay clear
    if n % 32:
        assert int(r["bits"][-1]) >> (n % 32) == 0
    tab.close()
    return r


def test_vertex_rows_with_repeats(gbp, env):
    """K candidates per tree vertex, the shape of newConfig: many candidates share a row"""
    name, t, o, T = env
    table = vertex_table(o, 300, seed=3)
    rng = np.random.default_rng(3)
    n = 20000
    idx = np.repeat(rng.integers(0, len(table), n // 50), 50).astype(np.int32)
    d = rng.integers(0, 2, n).astype(np.uint8)
    r = check(gbp, t, o, table, n, state_idx=idx, direction=d)
    assert 0 < r["n_valid"] < n


def test_direction_in_row_word(gbp, env):
    """gbp_sv_params.direction_in_row: rows and directions in ONE 4-byte word per candidate — same verdicts, rows and
    counters as the two-array form (every map class; fixed and adaptive step; directional sampling reads the direction too)"""
    name, t, o, T = env
    table = vertex_table(o, 700, seed=13)
    rng = np.random.default_rng(13)
    n = 12000
    idx = rng.integers(0, len(table), n).astype(np.int32)
    d = rng.integers(0, 2, n).astype(np.uint8)
    r = check(gbp, t, o, table, n, state_idx=idx, direction=d, packed=True)
    assert 0 < r["n_valid"] < n
    check(gbp, t, o, table, 3000, state_idx=idx[:3000], direction=d[:3000], packed=True, adaptive=True, states_valid=True)
    target = vertex_table(o, 1, seed=14)[0]
    check(gbp, t, o, table, 1000, state_idx=idx[:1000], direction=d[:1000], packed=True, target=target, thresh=0.15)
    # the flag excludes a separate direction array, and needs the row words
    tab = gbp.States(table)
    p = gbp.sv_params(1, 2, 3, direction_in_row=True)
    with pytest.raises(gbp.GbpError):
        t.sample_validate(tab, 100, p, idx[:100], d[:100])
    with pytest.raises(gbp.GbpError):
        t.sample_validate(tab, 100, p, None, None)
    tab.close()


def test_implicit_rows_with_offset_and_constant_direction(gbp, env):
    name, t, o, T = env
    table = vertex_table(o, 9000, seed=4)
    for direction0 in (0, 1):
        check(gbp, t, o, table, 8000, row0=517, direction0=direction0, idx0=123456789012)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 40, 63, 64, 65, 72, 73, 1023, 1025, 4100])
def test_ragged_sizes(gbp, env_sizes, n):
    name, t, o, T = env_sizes
    table = vertex_table(o, 4100, seed=6)
    rng = np.random.default_rng(n)
    check(gbp, t, o, table, n, direction=rng.integers(0, 2, n).astype(np.uint8))


def test_adaptive_step_and_tilted_normal(gbp, env):
    name, t, o, T = env
    table = vertex_table(o, 6000, seed=7)
    rng = np.random.default_rng(7)
    d = rng.integers(0, 2, 6000).astype(np.uint8)
    check(gbp, t, o, table, 6000, direction=d, adaptive=True, normal=(0.05, -0.02, 0.99))


def test_action_direction_sampling(gbp, env):
    """the fork's directional action sampling (planning_utils.cpp:379-391, :443-515) inside the kernel"""
    name, t, o, T = env
    table = vertex_table(o, 1500, seed=8)
    target = vertex_table(o, 1, seed=9)[0]
    rng = np.random.default_rng(8)
    d = rng.integers(0, 2, 1500).astype(np.uint8)
    for thresh in (0.15, 1.0):
        check(gbp, t, o, table, 1500, direction=d, target=target, thresh=thresh)


@pytest.mark.parametrize("adaptive", [False, True])
def test_start_states_promised_valid(gbp, env, adaptive):
    """start_states_valid = 1: the first sub-state of every candidate (the start state itself, valid by the caller's promise) is
    counted, not re-evaluated — verdicts, rows and the k / L / NaN-probe counters stay exactly the reference's.  Stance and
    flight times that skip whole phases included."""
    name, t, o, T = env
    table = vertex_table(o, 9000, seed=13)   # valid STANCE states: the promise holds
    rng = np.random.default_rng(13)
    d = rng.integers(0, 2, 9000).astype(np.uint8)
    check(gbp, t, o, table, 9000, direction=d, adaptive=adaptive, states_valid=True)
    check(gbp, t, o, table, 3000, direction0=1, adaptive=adaptive, states_valid=True, normal=(0.03, 0.01, 0.99))


def test_capped_valid_list(gbp, env):
    name, t, o, T = env
    table = vertex_table(o, 8000, seed=10)
    r = check(gbp, t, o, table, 8000, valid_cap=3)
    assert r["n_valid"] > 3 and len(r["index"]) == 3


def test_bad_arguments(gbp, env):
    name, t, o, T = env
    table = vertex_table(o, 64, seed=11)
    tab = gbp.States(table)
    p = gbp.sv_params(1, 1, 0)
    with pytest.raises(gbp.GbpError, match="outside the state table"):
        t.sample_validate(tab, 10, p, state_idx=np.array([0, 1, 2, 3, 64, 5, 6, 7, 8, 9], np.int32))
    with pytest.raises(gbp.GbpError, match="exceeds the state table"):
        t.sample_validate(tab, 65, p)
    p.direction0 = 3
    with pytest.raises(gbp.GbpError, match="direction0"):
        t.sample_validate(tab, 10, p)


def test_device_pointer_entry(gbp, env):
    """gbp_sample_validate_dev on torch device buffers and a torch stream: same bits, same rows"""
    torch = pytest.importorskip("torch")
    name, t, o, T = env
    n = 50000
    table = vertex_table(o, n, seed=12)
    rng = np.random.default_rng(12)
    idx = rng.permutation(n).astype(np.int32)
    d = rng.integers(0, 2, n).astype(np.uint8)
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        dt = torch.from_numpy(table).to(dev); di = torch.from_numpy(idx).to(dev); dd = torch.from_numpy(d).to(dev)
        bits = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device=dev)  # the call must clear the words itself
        fl = torch.empty(n, dtype=torch.uint8, device=dev)
        cap = 4096
        vi = torch.empty(cap, dtype=torch.int32, device=dev); sn = torch.empty((cap, 8), dtype=torch.float64, device=dev)
        tn = torch.empty(cap, dtype=torch.float64, device=dev); ac = torch.empty((cap, 10), dtype=torch.float64, device=dev)
        res = torch.zeros(8, dtype=torch.int64, device=dev)
        p = gbp.sv_params(21, 22, 7)
        for _ in range(2):  # twice on the same buffers: no state leaks from one call into the next
            t.sample_validate_dev(dt.data_ptr(), n, n, p, di.data_ptr(), dd.data_ptr(), bits.data_ptr(), fl.data_ptr(), cap, vi.data_ptr(),
                                  sn.data_ptr(), tn.data_ptr(), ac.data_ptr(), res.data_ptr(), st.cuda_stream)
    st.synchronize()
    s = table[idx]
    a = o.sample_actions(21, 22, 7, n)
    vo, fo, sno, tno, cnt = o.validate_pairs(s, a, d, nthreads=8)
    got = gbp.unpack_bits(bits.cpu().numpy().view(np.uint32), n)
    assert np.array_equal(got, vo)
    r = res.cpu().numpy()
    assert r[0] == vo.sum() and tuple(r[1:4]) == tuple(int(x) for x in cnt)
    want = np.flatnonzero(vo)
    m = min(len(want), cap)
    assert np.array_equal(vi.cpu().numpy()[:m], want[:m])
    assert_bits_equal(sn.cpu().numpy()[:m], sno[want[:m]], what="s_new (device entry)")
    assert_bits_equal(tn.cpu().numpy()[:m], tno[want[:m]], what="t_new (device entry)")
    assert_bits_equal(ac.cpu().numpy()[:m], a[want[:m]], what="action (device entry)")


@pytest.mark.parametrize("promise", [False, True])
def test_streamed_host_call_matches_chunked_and_oracle(gbp, env, monkeypatch, promise):
    """gbp_sample_validate on >= 2^20 candidates: ONE launch whose warps wait for the blocks of row numbers / directions the copy
    stream is still delivering (k_walk_sv_stream) against the launch-per-chunk form (GBP_SV_NO_STREAM) — verdict words, valid
    rows, work counters — and, on a slice, against the oracle."""
    name, t, o, T = env
    n = (1 << 20) + (1 << 19) + 12345  # three arrival blocks, the last one ragged
    table = vertex_table(o, 60000, seed=5)
    rng = np.random.default_rng(5)
    idx = rng.integers(0, len(table), n).astype(np.int32)
    d = rng.integers(0, 2, n).astype(np.uint8)
    tab = gbp.States(table)
    p = gbp.sv_params(31, 32, 1000, states_valid=promise)
    a = t.sample_validate(tab, n, p, state_idx=idx, direction=d, valid_cap=n, want_flags=False)
    monkeypatch.setenv("GBP_SV_NO_STREAM", "1")
    b = t.sample_validate(tab, n, p, state_idx=idx, direction=d, valid_cap=n, want_flags=False)
    monkeypatch.delenv("GBP_SV_NO_STREAM")
    assert np.array_equal(a["bits"], b["bits"]) and a["n_valid"] == b["n_valid"] and a["counters"] == b["counters"]
    assert np.array_equal(a["index"], b["index"])
    assert_bits_equal(a["s_new"], b["s_new"], what="s_new"); assert_bits_equal(a["t_new"], b["t_new"], what="t_new")
    assert_bits_equal(a["action"], b["action"], what="action")
    lo, m = (1 << 20) - 3000, 6000  # a slice across the first block boundary
    acts = o.sample_actions(31, 32, 1000 + lo, m)
    vo, fo, sno, tno, cnt = o.validate_pairs(table[idx[lo:lo + m]], acts, d[lo:lo + m], nthreads=8)
    assert np.array_equal(a["verdict"][lo:lo + m], vo)
    # and the one-word wire format (direction in bit 31 of the row word) through the same streamed launch
    pk = gbp.sv_params(31, 32, 1000, states_valid=promise, direction_in_row=True)
    c = t.sample_validate(tab, n, pk, state_idx=gbp.pack_rows(idx, d), valid_cap=n, want_flags=False)
    assert np.array_equal(a["bits"], c["bits"]) and a["n_valid"] == c["n_valid"] and a["counters"] == c["counters"]
    assert np.array_equal(a["index"], c["index"])
    assert_bits_equal(a["s_new"], c["s_new"], what="s_new (packed)"); assert_bits_equal(a["action"], c["action"], what="action (packed)")
