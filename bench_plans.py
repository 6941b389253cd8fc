"""Secondary metric of bench.py: solved plans/s (BASELINE configs[4]: 65,536 queries per GPU).

Q independent RRT-Connect queries per GPU on a seeded synthetic rough terrain, resident on the device
(gbp_plan_batch_dev: one warp per query), contiguous query ranges per rank, no inter-GPU traffic
except the final NCCL gather of the 80-byte per-query statistics records.  Reference-faithful
extend (K = 6 candidates, first valid decides), iteration budget instead of the wall clock.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
Q_PER_GPU = 65536  # BASELINE configs[4]'s batch on every GPU (weak scaling); 8192 per launch measured 30.8 k plans/s, 65536 33.8 k (fewer partial waves)
MAP_N, PITCH, BLOCK, STEP_H = 1024, 0.05, 8, 0.1
MAX_ITERS, MAX_VERTS, K_CAND = 2000, 512, 6


def rough_terrain(seed=3):
    rng = np.random.default_rng(seed)
    ax = np.arange(MAP_N) * PITCH
    b = rng.uniform(0, STEP_H, (MAP_N // BLOCK + 1, MAP_N // BLOCK + 1))
    z = np.kron(b, np.ones((BLOCK, BLOCK)))[:MAP_N, :MAP_N]
    return ax, ax.copy(), z.astype(np.float32).astype(np.float64)


def make_queries(valid_states_fn, sample_states_fn, nq, seed, stream):
    """Start/goal pairs: valid STANCE poses 3-5 m apart, velocity 0.5 m/s along the start->goal line,
    pitch 0 (both re-validated).  Deterministic in (seed, stream)."""
    starts, goals = [], []
    idx0 = 0
    rng = np.random.default_rng(seed * 1000003 + stream)
    while len(starts) < nq:
        m = 4 * nq
        q = sample_states_fn(seed, stream, idx0, m); idx0 += m
        q[:, 3:8] = 0.0
        ang = rng.uniform(0, 2 * np.pi, m); dist = rng.uniform(3.0, 5.0, m)
        g = q.copy()
        g[:, 0] += dist * np.cos(ang); g[:, 1] += dist * np.sin(ang)
        q[:, 3] = g[:, 3] = 0.5 * np.cos(ang); q[:, 4] = g[:, 4] = 0.5 * np.sin(ang)
        lim = MAP_N * PITCH - 1.0
        ok = (g[:, 0] > 1.0) & (g[:, 0] < lim) & (g[:, 1] > 1.0) & (g[:, 1] < lim) & (q[:, 0] > 1.0) & (q[:, 0] < lim) & (q[:, 1] > 1.0) & (q[:, 1] < lim)
        q, g = q[ok], g[ok]
        vq = valid_states_fn(q); q, g = q[vq == 1], g[vq == 1]
        starts.extend(q); goals.extend(g)
    return np.array(starts[:nq]), np.array(goals[:nq])


def run_rough_k4096(gbp, torch, dev, nq=7104, iters=40):
    """BASELINE configs[1]: RRT-Connect on the reference's data/rough_terrain (committed as
    tests/golden/terrain_rough_terrain.npz), 4096 candidate actions per extend (closest valid), start (0,0) ->
    goal (8,0) at body height 0.375 m; nq independent searches (distinct Philox streams; 7104 = two
    waves of the planner's 24 resident warps per SM on 148 SMs)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "terrain_rough_terrain.npz"))
    t = gbp.Terrain(d["x"], d["y"], d["z"], d["dx"], d["dy"], d["dz"])
    h, _ = t.ground_height([0.0, 8.0], [0.0, 0.0])
    start = np.array([0, 0, h[0] + 0.375, 1, 0, 0, 0, 0.0]); goal = np.array([8, 0, h[1] + 0.375, 1, 0, 0, 0, 0.0])
    s = torch.from_numpy(np.repeat(start[None], nq, 0)).to(dev); g = torch.from_numpy(np.repeat(goal[None], nq, 0)).to(dev)
    P = gbp.PlanParams(4096, 1, iters, 256, 0, 0, 0)
    dstats = torch.zeros(nq * 80, dtype=torch.uint8, device=dev)
    cur = torch.cuda.current_stream().cuda_stream
    t.plan_batch_dev(nq, s.data_ptr(), g.data_ptr(), 1, 0, P, dstats.data_ptr(), cur)  # warm-up (also sizes the tree arena)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t.plan_batch_dev(nq, s.data_ptr(), g.data_ptr(), 1, 0, P, dstats.data_ptr(), cur)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    st = dstats.cpu().numpy().view(gbp.PLAN_STATS_DTYPE)
    # ONE search driven extend by extend through the host-pointer call (gbp_extend: nearest neighbour, 4096 candidates
    # sampled and validated in-kernel, selection and append on the device; 64 B in, 16 B out per call)
    tree = gbp.Tree(8192, start)
    targets = t.sample_states(5, 77, 0, 4000)
    targets = targets[t.valid_states(targets, gbp.STANCE)[0] == 1][:600]
    for i in range(20):
        tree.extend(t, targets[i], gbp.FORWARD, 4096, 1, 5, 78, i * 4096)
    t0 = time.perf_counter()
    added = 0
    for i in range(20, len(targets)):
        stt, nid, chk = tree.extend(t, targets[i], gbp.FORWARD, 4096, 1, 5, 78, i * 4096)
        added += int(stt != gbp.TRAPPED)
    dt_single = time.perf_counter() - t0
    n_ext = len(targets) - 20
    single = {"extends_per_s": n_ext / dt_single, "validated_actions_per_s": n_ext * 4096 / dt_single, "us_per_extend": dt_single / n_ext * 1e6,
              "extends": n_ext, "vertices_added": added, "api": "gbp_extend (host pointers, one synchronous call per extend)"}
    return {"workload": f"{nq} searches on data/rough_terrain, (0,0)->(8,0), K=4096 closest-valid candidates per extend, {iters} iterations",
            "single_search": single,
            "validated_actions_per_s": float(st["pair_checks"].sum() / secs), "extends_per_s": float(st["nn_queries"].sum() / 2 / secs),
            "solved": int(st["solved"].sum()), "solved_plans_per_s": float(st["solved"].sum() / secs), "seconds": secs,
            "note": "the unmodified CPU reference completes 0 plans in 120 s on this query (BASELINE.md row 9)"}


def run_slope_config0(gbp, rounds=8):
    """BASELINE configs[0]: RRT-Connect on the reference's data/slope (tests/golden/terrain_slope.npz), (0,0) -> (8,0) at body
    height 0.30 m (the fork's 0.375 m start pose is invalid on this map), default parameters.
    Time to a first solution through the host-pointer C ABI: one gbp_plan_batch launch of 3552 attempts at the same query
    (distinct Philox streams, 8000 iterations / 2048 vertices per tree) that stops when the first attempt has solved —
    what RRTConnectClass::buildRRTConnect of the drop-in does in its first anytime round."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "terrain_slope.npz"))
    t = gbp.Terrain(d["x"], d["y"], d["z"], d["dx"], d["dy"], d["dz"])
    h, _ = t.ground_height([0.0, 8.0], [0.0, 0.0])
    start = np.array([0, 0, h[0] + 0.30, 1, 0, 0, 0, 0.0]); goal = np.array([8, 0, h[1] + 0.30, 1, 0, 0, 0, 0.0])
    nq = 3552
    S, G = np.repeat(start[None], nq, 0), np.repeat(goal[None], nq, 0)
    P = gbp.PlanParams(6, 0, 8000, 2048, 0, 0, 0, 1)
    t.plan_batch(S[:64], G[:64], 1, 0, gbp.PlanParams(6, 0, 10, 2048, 0, 0, 0, 0))  # warm-up: sizes the tree arena
    t.plan_batch(S, G, 1, 1 << 40, gbp.PlanParams(6, 0, 10, 2048, 0, 0, 0, 0))
    times, solved, lengths, launches, q0 = [], 0, [], 0, nq
    for r in range(rounds):
        t0 = time.perf_counter()
        for attempt_round in range(60):  # anytime rounds of one call: fresh Philox streams until an attempt solves
            st = t.plan_batch(S, G, 1, q0, P)
            q0 += nq
            launches += 1
            ok = st["solved"] == 1
            if ok.any():
                solved += 1
                lengths.append(float(st["path_length"][ok].min()))
                break
        times.append(time.perf_counter() - t0)
    return {"workload": f"data/slope, (0,0)->(8,0), body 0.30 m, K=6 first-valid; {rounds} independent calls; a call launches rounds of {nq} "
                        "device-resident attempts (each round ends with its first solution) until one solves (host-pointer C ABI, wall clock)",
            "calls": rounds, "calls_solved": solved, "rounds_launched": launches,
            "first_solution_s": {"mean": float(np.mean(times)), "min": float(np.min(times)), "max": float(np.max(times))},
            "raw_path_length_m_mean": float(np.mean(lengths)) if lengths else None,
            "reference_note": "unmodified reference, 1 core (BASELINE.md row 8): first solutions after 4.6 s and 6.7 s, 2 plans in 93 s"}


def run(gbp, torch, dist, dev, rank, world, q_per_gpu=Q_PER_GPU, cpu_seconds=8.0, want_cpu=True):
    x, y, z = rough_terrain()
    t = gbp.Terrain(x, y, z)
    seed, stream = 11, 7000 + rank
    # goal height: body height of the start above the goal's own ground
    def valid(s):
        return t.valid_states(s, gbp.STANCE)[0]
    s, g = make_queries(valid, t.sample_states, 2 * q_per_gpu, seed, stream)
    hs, _ = t.ground_height(s[:, 0], s[:, 1]); hg, _ = t.ground_height(g[:, 0], g[:, 1])
    g[:, 2] = s[:, 2] - hs + hg
    vg = valid(g)
    s, g = s[vg == 1][:q_per_gpu], g[vg == 1][:q_per_gpu]
    nq = len(s)
    P = gbp.PlanParams(K_CAND, 0, MAX_ITERS, MAX_VERTS, 0, 0, 0)
    ds = torch.from_numpy(s).to(dev); dg = torch.from_numpy(g).to(dev)
    dstats = torch.zeros(nq * 80, dtype=torch.uint8, device=dev)
    cur = torch.cuda.current_stream().cuda_stream
    query0 = rank * q_per_gpu
    t.plan_batch_dev(nq, ds.data_ptr(), dg.data_ptr(), seed, query0, P, dstats.data_ptr(), cur)  # warm-up (also sizes the tree arena)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t.plan_batch_dev(nq, ds.data_ptr(), dg.data_ptr(), seed, query0, P, dstats.data_ptr(), cur)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        gathered = [torch.empty_like(dstats) for _ in range(world)] if rank == 0 else None
        dist.gather(dstats, gathered, dst=0)  # the final NCCL gather of plan statistics (80 B per query)
        allstats = torch.cat(gathered).cpu().numpy() if rank == 0 else None
    else:
        allstats = dstats.cpu().numpy()
    if rank != 0:
        return None
    st = allstats.view(gbp.PLAN_STATS_DTYPE)
    secs = float(ms.item()) * 1e-3
    out = {"workload": f"{nq} RRT-Connect queries per GPU (BASELINE configs[4]), synthetic rough terrain {MAP_N}x{MAP_N} @ {PITCH} m, "
                       f"blocks {BLOCK * PITCH:.1f} m, steps U(0,{STEP_H}) m; start/goal 3-5 m apart; K={K_CAND} first-valid, "
                       f"budget {MAX_ITERS} iterations / {MAX_VERTS} vertices per tree",
           "queries": int(len(st)), "solved": int(st["solved"].sum()), "solved_plans_per_s": float(st["solved"].sum() / secs),
           "queries_per_s": float(len(st) / secs), "validated_actions_per_s": float(st["pair_checks"].sum() / secs),
           "nn_queries_per_s": float(st["nn_queries"].sum() / secs), "seconds": secs,
           "mean_path_length_m": float(st["path_length"][st["solved"] == 1].mean()) if st["solved"].any() else None,
           "mean_iters": float(st["iters"].mean()), "stats_gather_bytes": int(allstats.nbytes)}
    out["rough_k4096"] = run_rough_k4096(gbp, torch, dev) if world == 1 or rank == 0 else None
    out["slope_config0"] = run_slope_config0(gbp)
    # BASELINE configs[2]: RRT*-Connect (choose parent + near-set rewiring, delta = 3 m, rrt_star_connect.cpp:12-75) with
    # postProcessPath, on the first queries of the same set; iteration budget instead of the wall-clock budget
    ns, star_iters = min(nq, 2048), 400
    Ps = gbp.PlanParams(K_CAND, 0, star_iters, MAX_VERTS, 0, 1, 1)
    Pc2 = gbp.PlanParams(K_CAND, 0, star_iters, MAX_VERTS, 0, 0, 1)
    res = {}
    for name, PP in (("rrt_star_connect", Ps), ("rrt_connect_same_budget", Pc2)):
        t.plan_batch_dev(ns, ds.data_ptr(), dg.data_ptr(), seed, query0, PP, dstats.data_ptr(), cur)
        torch.cuda.synchronize()
        e0.record()
        t.plan_batch_dev(ns, ds.data_ptr(), dg.data_ptr(), seed, query0, PP, dstats.data_ptr(), cur)
        e1.record()
        torch.cuda.synchronize()
        ss = dstats.cpu().numpy().view(gbp.PLAN_STATS_DTYPE)[:ns]
        sec = e0.elapsed_time(e1) * 1e-3
        ok = ss["solved"] == 1
        res[name] = {"solved": int(ok.sum()), "solved_plans_per_s": float(ok.sum() / sec), "seconds": sec,
                     "mean_path_length_m": float(ss["path_length"][ok].mean()) if ok.any() else None,
                     "validated_actions_per_s": float(ss["pair_checks"].sum() / sec), "mean_vertices": float((ss["nv_a"] + ss["nv_b"]).mean())}
    res["workload"] = (f"{ns} queries of the same set, {star_iters} iterations, post-processed paths; RRT*-Connect (delta = 3 m rewiring) against "
                       "RRT-Connect at the same budget")
    out["rrt_star"] = res
    if want_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle as po
        o = po.Oracle(po.Terrain(x, y, z))
        cores = os.cpu_count() or 1
        Pc = po.PlanParams(K_CAND, 0, MAX_ITERS, MAX_VERTS, 0, 0, 0)
        m = min(nq, 4 * cores)
        t0 = time.perf_counter(); cst = o.plan_batch(s[:m], g[:m], seed, 0, Pc, cores); dt = time.perf_counter() - t0
        m2 = int(min(nq, max(m, m / dt * cpu_seconds)))
        if m2 > m:
            t0 = time.perf_counter(); cst = o.plan_batch(s[:m2], g[:m2], seed, 0, Pc, cores); dt = time.perf_counter() - t0
            m = m2
        gst = st[:m]
        same = all(np.array_equal(cst[k], gst[k]) for k in ("solved", "iters", "nv_a", "nv_b", "pair_checks", "nn_queries", "path_states"))
        out["cpu_baseline"] = {"kind": "port", "cores": cores, "queries": m, "solved": int(cst["solved"].sum()),
                               "solved_plans_per_s": float(cst["solved"].sum() / dt), "queries_per_s": m / dt,
                               "validated_actions_per_s": float(cst["pair_checks"].sum() / dt), "seconds": dt,
                               "note": "iteration-budgeted oracle planner (reference L0-L2 restated, same Philox stream); the "
                                       "reference's own loops are wall-clock driven and not reproducible",
                               "gpu_stats_identical_on_sample": bool(same),
                               "path_length_max_rel_diff": float(np.max(np.abs(cst["path_length"] - gst["path_length"]) /
                                                                        np.maximum(1e-300, np.abs(cst["path_length"]))) if m else 0.0)}
    return out
