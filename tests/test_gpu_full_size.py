"""Size-independent properties at BASELINE's FULL size (configs[3]: 16,777,216 candidates on the synthetic 4096 x 4096
map — exactly bench.py's batch, built by bench.device_batch).  The oracle cannot walk 16.7 M candidates in seconds, so
the whole batch is checked through properties the path must have at any size, all bit for bit:

* the shipped three-kernel path (mixed-precision walk + fp64 redo + exact outputs, variant 0) equals the independent
  fp64 thread-per-action kernel (variant 1) on every candidate: verdicts, s_new, t_new and the k / L work counters;
* determinism (two launches), split invariance (one call == ragged chunks) and permutation equivariance;
* a fully valid FORWARD candidate's s_new is its landing state applyFlight(applyStance(s, a, t_s), t_f), t_new = t_s + t_f;
* the head of the batch equals the oracle restatement (and, when built, the unmodified reference) candidate by candidate.
"""
import numpy as np
import pytest

import pyoracle as po
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    import __graft_entry__ as entry
    entry.build()
    import torch
    import bench
    import global_body_planner_b200 as gbp
    from global_body_planner_b200 import capi
    assert gbp.device_count() > 0, "no CUDA device: the GPU tests must not pass on a fallback"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    gbp.set_device(0)
    x, y, z = bench.synthetic_map()
    t = gbp.Terrain(x, y, z)
    assert t.flags()["mixed_precision"]
    n = bench.N_CAND
    s, a, d = bench.device_batch(torch, capi, t, x, y, n, 1, 100, dev)
    out = run(torch, t, n, s, a, d, 0)
    cnt = t.validate_counters()
    return dict(torch=torch, gbp=gbp, t=t, n=n, s=s, a=a, d=d, out=out, cnt=cnt, map=(x, y, z), bench=bench)


def run(torch, t, n, s, a, d, variant, first=0):
    """one device-pointer call over candidates [first, first + n) of the arrays"""
    dev = s.device
    v = torch.zeros(n, dtype=torch.uint8, device=dev); f = torch.zeros_like(v)
    sn = torch.zeros((n, 8), dtype=torch.float64, device=dev); tn = torch.zeros(n, dtype=torch.float64, device=dev)
    t.validate_pairs_dev(n, s.data_ptr() + 64 * first, a.data_ptr() + 80 * first, d.data_ptr() + first, 0, variant,
                         v.data_ptr(), f.data_ptr(), sn.data_ptr(), tn.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return v, f, sn, tn


def same(torch, A, B):
    """verdicts, the out-of-grid flag, s_new and t_new bit for bit (int64 views: NaN-safe, -0 != +0)"""
    return (torch.equal(A[0], B[0]) and torch.equal(A[1] & 2, B[1] & 2) and torch.equal(A[2].view(torch.int64), B[2].view(torch.int64))
            and torch.equal(A[3].view(torch.int64), B[3].view(torch.int64)))


def test_shipped_path_equals_fp64_kernel_on_every_candidate(full):
    torch, t, n = full["torch"], full["t"], full["n"]
    ref = run(torch, t, n, full["s"], full["a"], full["d"], 1)
    c1 = t.validate_counters()
    assert same(torch, full["out"], ref), "mixed-precision walk path differs from the fp64 thread-per-action kernel"
    c0 = full["cnt"]
    assert (c0["substates"], c0["lookups"], c0["nanprobes"], c0["valid"]) == (c1["substates"], c1["lookups"], c1["nanprobes"], c1["valid"])
    assert c0["valid"] == int(full["out"][0].sum().item()) and 0 < c0["valid"] < n
    assert 2 * n <= c0["substates"] <= 19 * n


def test_determinism_and_split_invariance(full):
    torch, t, n = full["torch"], full["t"], full["n"]
    assert same(torch, full["out"], run(torch, t, n, full["s"], full["a"], full["d"], 0))
    cuts = [0, 5_000_016, 5_000_064, 11_111_120, n]  # ragged for every tile size in use, 16-candidate aligned
    parts = [run(torch, t, cuts[i + 1] - cuts[i], full["s"], full["a"], full["d"], 0, first=cuts[i]) for i in range(len(cuts) - 1)]
    glued = tuple(torch.cat([p[k] for p in parts]) for k in range(4))
    assert same(torch, full["out"], glued), "chunked calls differ from the single call"


def test_permutation_equivariance(full):
    torch, t, n = full["torch"], full["t"], full["n"]
    g = torch.Generator(device=full["s"].device); g.manual_seed(5)
    perm = torch.randperm(n, device=full["s"].device, generator=g)
    out = run(torch, t, n, full["s"][perm].contiguous(), full["a"][perm].contiguous(), full["d"][perm].contiguous(), 0)
    want = tuple(o[perm] for o in full["out"])
    assert same(torch, want, out), "results depend on the position of a candidate in the batch"


def test_valid_forward_candidates_land_on_the_propagated_state(full):
    torch, gbp, n = full["torch"], full["gbp"], full["n"]
    v, f, sn, tn = full["out"]
    idx = torch.nonzero((v == 1) & (full["d"] == 0))[:, 0][:200_000]
    assert len(idx) > 1000
    s, a = full["s"][idx].cpu().numpy(), full["a"][idx].cpu().numpy()
    o = po.Oracle()
    land = o.apply_flight(o.apply_stance(s, a, a[:, 6]), a[:, 7])
    assert_bits_equal(sn[idx].cpu().numpy(), land, what="s_new of fully valid FORWARD candidates")
    assert np.array_equal(tn[idx].cpu().numpy(), a[:, 6] + a[:, 7])


def test_head_of_the_batch_equals_oracle_and_reference(full):
    torch, t, bench = full["torch"], full["t"], full["bench"]
    m = 1 << 16
    x, y, z = full["map"]
    T = po.Terrain(x, y, z)
    o = po.Oracle(T)
    s, a, d = (full[k][:m].cpu().numpy() for k in ("s", "a", "d"))
    hs, ha, hd = bench.cpu_candidates(m, 1, 100)  # the host regeneration of the same Philox cells
    assert np.array_equal(s.view(np.uint64), hs.view(np.uint64)) and np.array_equal(a.view(np.uint64), ha.view(np.uint64)) and np.array_equal(d, hd)
    v, f, sn, tn = (q[:m].cpu().numpy() for q in full["out"])
    vo, fo, sno, tno, _ = o.validate_pairs(s, a, d, nthreads=8)
    assert np.array_equal(v, vo) and np.array_equal(f & 2, fo & 2)
    assert_bits_equal(sn, sno, what="s_new vs oracle")
    assert_bits_equal(tn, tno, what="t_new vs oracle")
    if po.Ref.available():
        k = 1 << 13  # the unmodified reference copies both 4096-entry axes per isValidState (planning_utils.cpp:568-569)
        vr = np.asarray(po.Ref(T).validate_pairs(s[:k], a[:k], d[:k], False, 8)[0])
        ingrid = (f[:k] & 2) == 0
        assert np.array_equal(vr[ingrid], v[:k][ingrid])
