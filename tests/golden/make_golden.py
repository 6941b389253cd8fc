"""Mint tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libgbp_ref.so).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The GPU box has no /root/reference; its tests compare the CUDA path and the oracle restatement with
these committed vectors.  Inputs are drawn from the Philox stream spec (oracle restatement) so they
are reproducible; every OUTPUT stored here comes from the reference build, except `ingrid`, which
is the oracle's out-of-grid flag (the reference has undefined behaviour there)
and `unique`, the oracle's "minimum attained once" flag for nearest-neighbour queries.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyoracle as po  # noqa: E402

N_PAIRS = 3000


def synth_nan_terrain():
    """256x256 @ 0.05 m, rolling ground + seeded fp32 noise + NaN holes (exercises heightIsNan)."""
    rng = np.random.default_rng(7)
    n, pitch = 256, 0.05
    ax = np.arange(n) * pitch - 1.0
    z = 0.05 * np.sin(0.7 * ax)[:, None] * np.cos(0.5 * ax)[None, :] + rng.normal(0, 0.01, (n, n))
    z = z.astype(np.float32).astype(np.float64)
    for _ in range(40):
        i, j = rng.integers(8, n - 8, 2)
        z[i:i + rng.integers(1, 4), j:j + rng.integers(1, 4)] = np.nan
    dz = np.ones_like(z)
    dx = rng.normal(0, 0.05, (n, n)); dy = rng.normal(0, 0.05, (n, n))
    return po.Terrain(ax, ax.copy(), z, dx, dy, dz)


def synth_mixed_terrain():
    """384x384 @ 0.05 m, rolling ground + seeded fp32 noise, no NaN, uniform axes: the terrain class served by the
    mixed-precision walk kernel (k_walk_mixed + fp64 redo pass) and its texture-gather fetch path."""
    rng = np.random.default_rng(17)
    n, pitch = 384, 0.05
    ax = np.arange(n) * pitch - 2.0
    z = 0.06 * np.sin(0.9 * ax)[:, None] * np.cos(0.6 * ax)[None, :] + rng.normal(0, 0.008, (n, n))
    z = z.astype(np.float32).astype(np.float64)
    dz = np.ones_like(z)
    dx = rng.normal(0, 0.05, (n, n)); dy = rng.normal(0, 0.05, (n, n))
    return po.Terrain(ax, ax.copy(), z, dx, dy, dz)


def mint(name, T, seed):
    o, r = po.Oracle(T), po.Ref(T)
    out = {}
    rng = np.random.default_rng(seed)
    # terrain probes strictly inside the grid
    m = 2000
    px = rng.uniform(T.x[0], T.x[-1] - 1e-9, m); py = rng.uniform(T.y[0], T.y[-1] - 1e-9, m)
    px[:8] = T.x[:8]; py[:8] = T.y[:8]  # exact grid lines
    out.update(probe_x=px, probe_y=py, probe_h=r.ground_height(px, py), probe_nan=r.height_is_nan(px, py),
               probe_normal=r.surface_normal(px, py))
    # random states through the stream spec; keep the mix of valid / invalid
    q = o.sample_states(seed, 3, 0, 6000)
    _, fl = o.valid_states(q, po.STANCE)
    q = q[(fl & po.FLAG_OOG) == 0][:3000]
    out.update(vs_states=q, vs_stance=r.valid_states(q, po.STANCE), vs_flight=r.valid_states(q, po.FLIGHT))
    # candidate pairs from valid stance states
    big = o.sample_states(seed, 5, 0, 40000)
    v, _ = o.valid_states(big, po.STANCE)
    s = big[v == 1][:N_PAIRS]
    normal = r.surface_normal(s[:1, 0], s[:1, 1])[0] if not name.startswith("synth") else np.array([0.03, -0.02, 0.99])
    a = o.sample_actions(seed, 9, 0, len(s), normal)
    a[: len(s) // 8, 7] = 0.0  # some pure-stance primitives
    g0 = 2 * len(s) // 3       # last third: gentle primitives so that the fully-valid branch is well covered
    a[g0:, [0, 1, 2, 3, 4, 5, 8, 9]] = rng.normal(0, 1.0, (len(s) - g0, 8))
    a[g0:, 7] = rng.uniform(0, 0.15, len(s) - g0)
    d = (np.arange(len(s)) % 2).astype(np.uint8)
    out.update(pair_states=s, pair_actions=a, pair_dir=d, pair_normal=normal)
    for adaptive in (0, 1):
        vr, snr, tnr = r.validate_pairs(s, a, d, adaptive=bool(adaptive))
        _, fo, _, _, _ = o.validate_pairs(s, a, d, adaptive=bool(adaptive))
        out[f"pair_verdict_{adaptive}"] = vr
        out[f"pair_snew_{adaptive}"] = snr   # NaN rows/entries = the reference left them unwritten
        out[f"pair_tnew_{adaptive}"] = tnr
        out[f"pair_ingrid_{adaptive}"] = ((fo & po.FLAG_OOG) == 0).astype(np.uint8)
    # propagation
    tt = rng.uniform(0, 0.3, len(s))
    out.update(prop_t=tt, prop_stance=r.apply_stance(s, a, tt), prop_flight=r.apply_flight(s, tt - 0.1),
               prop_stance_rev=r.apply_stance_reverse(s, a, tt))
    # attemptConnect between valid states (and isValidAction on the resulting actions)
    k = min(len(s) // 2, 1000)
    se, sg = s[:k], s[k:2 * k].copy()
    sg[: k // 2, :3] = se[: k // 2, :3] + rng.normal(0, 0.25, (k // 2, 3)) * np.array([1, 1, 0.1])  # nearby targets
    cd = (np.arange(k) % 2).astype(np.uint8)
    st, sn, an = r.attempt_connect(se, sg, cd)
    _, _, _, fo = o.attempt_connect(se, sg, cd)
    out.update(con_existing=se, con_target=sg, con_dir=cd, con_status=st, con_snew=sn, con_anew=an,
               con_ingrid=((fo & po.FLAG_OOG) == 0).astype(np.uint8))
    acts = np.concatenate([an[~np.isnan(an).any(axis=1)], a[:500]])
    acts[::7, 2] -= 30.0
    out.update(va_actions=acts, va_verdict=r.is_valid_action(acts))
    # tree queries
    verts = s[:1500]
    qs = q[:200]
    idx, dist = r.nearest(verts, qs)
    _, _, uniq = o.nearest(verts, qs)
    out.update(nn_verts=verts, nn_q=qs, nn_idx=idx, nn_dist=dist, nn_unique=uniq.astype(np.uint8))
    near = [np.sort(r.near(verts, qq, 3.0)) for qq in qs[:20]]
    out.update(near_counts=np.array([len(x) for x in near]), near_ids=np.concatenate(near).astype(np.int32))
    out.update(dist_pose=r.distance(verts[:500], verts[500:1000], 0), dist_state=r.distance(verts[:500], verts[500:1000], 1),
               dist_yaw=r.distance(verts[:500], verts[500:1000], 2))
    parent = np.array([-1] + [int(rng.integers(0, i)) for i in range(1, 300)], np.int32)
    g, y = r.tree_gy(verts[:300], parent)
    out.update(gy_parent=parent, gy_g=g, gy_y=y)
    f3 = rng.normal(0, 100, (200, 3)); n3 = rng.normal(0, 0.2, (200, 3)) + np.array([0, 0, 1.0]); n3[:5] = [0, 0, 1]
    out.update(grf_n=n3, grf_f=f3, grf_out=r.rotate_grf(n3, f3))
    # plan output: getInterpPath + calculateMaxCurvature over a 40-primitive sequence (half of them pure stance)
    i0 = len(s) // 8 - 20
    ps, pa = s[i0:i0 + 41], a[i0:i0 + 40]
    ips, ipt, ipp = r.interp_path(ps, pa, 0.05)
    out.update(interp_in_states=ps, interp_in_actions=pa, interp_states=ips, interp_t=ipt, interp_phase=ipp,
               interp_max_curvature=np.array(r.max_curvature(ips)))
    np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), **out)
    r.close()
    print(name, {k2: getattr(v2, "shape", None) for k2, v2 in list(out.items())[:4]}, "valid pairs:",
          int(out["pair_verdict_0"].sum()), "connect status hist:", np.bincount(st, minlength=3))


if __name__ == "__main__":
    if not po.Ref.available():
        po.build()
    for name in ("rough_terrain", "slope"):
        T = po.Terrain.from_reference_csv(os.path.join(po.REFERENCE_ROOT, "data", name))
        T.save_npz(os.path.join(HERE, f"terrain_{name}.npz"))
        mint(name, T, seed=11 if name == "slope" else 12)
    T = synth_nan_terrain()
    np.savez_compressed(os.path.join(HERE, "terrain_synth_nan.npz"), x=T.x, y=T.y, z=T.z.astype(np.float32),
                        dx=T.dx.astype(np.float32), dy=T.dy.astype(np.float32), dz=T.dz.astype(np.float32))
    T = po.Terrain.from_npz(os.path.join(HERE, "terrain_synth_nan.npz"))  # exactly what the tests will load
    mint("synth_nan", T, seed=13)
    T = synth_mixed_terrain()
    np.savez_compressed(os.path.join(HERE, "terrain_synth_mixed.npz"), x=T.x, y=T.y, z=T.z.astype(np.float32),
                        dx=T.dx.astype(np.float32), dy=T.dy.astype(np.float32), dz=T.dz.astype(np.float32))
    T = po.Terrain.from_npz(os.path.join(HERE, "terrain_synth_mixed.npz"))
    mint("synth_mixed", T, seed=14)
