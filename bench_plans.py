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


def make_queries(valid_states_fn, sample_states_fn, nq, seed, stream, dist_lo=4.0, dist_hi=8.0):
    """Start/goal pairs: valid STANCE poses dist_lo..dist_hi m apart (4-8 m), velocity 0.5 m/s along the
    start->goal line, pitch 0 (both re-validated).  Deterministic in (seed, stream)."""
    starts, goals = [], []
    idx0 = 0
    rng = np.random.default_rng(seed * 1000003 + stream)
    while len(starts) < nq:
        m = 4 * nq
        q = sample_states_fn(seed, stream, idx0, m); idx0 += m
        q[:, 3:8] = 0.0
        ang = rng.uniform(0, 2 * np.pi, m); dist = rng.uniform(dist_lo, dist_hi, m)
        g = q.copy()
        g[:, 0] += dist * np.cos(ang); g[:, 1] += dist * np.sin(ang)
        q[:, 3] = g[:, 3] = 0.5 * np.cos(ang); q[:, 4] = g[:, 4] = 0.5 * np.sin(ang)
        lim = MAP_N * PITCH - 1.0
        ok = (g[:, 0] > 1.0) & (g[:, 0] < lim) & (g[:, 1] > 1.0) & (g[:, 1] < lim) & (q[:, 0] > 1.0) & (q[:, 0] < lim) & (q[:, 1] > 1.0) & (q[:, 1] < lim)
        q, g = q[ok], g[ok]
        vq = valid_states_fn(q); q, g = q[vq == 1], g[vq == 1]
        starts.extend(q); goals.extend(g)
    return np.array(starts[:nq]), np.array(goals[:nq])


def _ref_worker(name, algorithm, start, goal, max_time, adaptive, q):
    """one call of the UNMODIFIED buildRRTConnect / buildRRTStarConnect (oracle/_ref/libgbp_ref.so), in its own process"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    d = np.load(os.path.join(ROOT, "tests", "golden", f"terrain_{name}.npz"))
    r = po.Ref(po.Terrain(d["x"], d["y"], d["z"], d["dx"], d["dy"], d["dz"]))
    t0 = time.perf_counter()
    st, ss, aa = r.plan(algorithm, start, goal, max_time, adaptive)
    st["wall_s"] = time.perf_counter() - t0
    q.put(st)


def reference_cpu_planner(name, algorithm, start, goal, max_time, budget_s, procs=None):
    """The reference's CPU planner timed in the same run: `procs` independent processes each make ONE call
    of the unmodified buildRRTConnect (algorithm 0) / buildRRTStarConnect (1) exactly as callPlanner does
    (global_body_planner.cpp:113-124) with replan_time_limit = max_time; rand() and the clock-seeded engines make every call
    different.  The reference has no iteration budget and a 4000 s give-up time (rrt_connect.h:119-120): calls still
    running after budget_s are killed and counted as unsolved."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    if not po.Ref.available():
        return {"unavailable": "oracle/_ref/libgbp_ref.so not built"}
    procs = procs or min(os.cpu_count() or 1, 16)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_ref_worker, args=(name, algorithm, start, goal, max_time, 0, q)) for _ in range(procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    done = []
    while time.perf_counter() - t0 < budget_s and len(done) < procs:
        try:
            done.append(q.get(timeout=0.25))
        except Exception:
            pass
    wall = time.perf_counter() - t0
    for p in ps:
        if p.is_alive():
            p.terminate()
        p.join(timeout=2)
    solved = [d for d in done if d["n_states"] > 0]
    out = {"kind": "reference", "api": "buildRRTStarConnect" if algorithm else "buildRRTConnect", "processes": procs, "calls_finished": len(done),
           "calls_killed_at_budget": procs - len(done), "budget_s": budget_s, "replan_time_limit_s": max_time, "wall_s": wall,
           "solved_plans": len(solved), "solved_plans_per_s": len(solved) / wall}
    if solved:
        out.update(time_to_first_solution_s={"mean": float(np.mean([d["time_to_first"] for d in solved])), "min": float(np.min([d["time_to_first"] for d in solved]))},
                   plan_time_s_mean=float(np.mean([d["plan_time"] for d in solved])), path_cost_m_mean=float(np.mean([d["cost"] for d in solved])),
                   validated_actions_per_s_per_core=float(np.mean([d["pair_checks"] / d["wall_s"] for d in solved])),
                   nn_queries_per_s_per_core=float(np.mean([d["nn_queries"] / d["wall_s"] for d in solved])))
    return out


def shipped_query(gbp, name, body):
    d = np.load(os.path.join(ROOT, "tests", "golden", f"terrain_{name}.npz"))
    t = gbp.Terrain(d["x"], d["y"], d["z"], d["dx"], d["dy"], d["dz"])
    h, _ = t.ground_height([0.0, 8.0], [0.0, 0.0])
    return t, np.array([0, 0, h[0] + body, 1, 0, 0, 0, 0.0]), np.array([8, 0, h[1] + body, 1, 0, 0, 0, 0.0])


def first_solution_rounds(t, start, goal, P, nq, rounds, max_launches=60):
    """`rounds` independent calls; a call launches rounds of nq device-resident attempts at the same query (distinct Philox
    streams, stop_after_solved = 1) until one solves — what build* of the drop-in does in its first anytime round"""
    S, G = np.repeat(start[None], nq, 0), np.repeat(goal[None], nq, 0)
    times, solved, costs, launches, q0 = [], 0, [], 0, nq
    for r in range(rounds):
        t0 = time.perf_counter()
        for _ in range(max_launches):
            st = t.plan_batch(S, G, 1, q0, P)
            q0 += nq
            launches += 1
            ok = st["solved"] == 1
            if ok.any():
                solved += 1
                costs.append(float(st["path_cost"][ok].min()))
                break
        times.append(time.perf_counter() - t0)
    return {"calls": rounds, "calls_solved": solved, "launches": launches, "attempts_per_launch": nq,
            "time_to_first_solution_s": {"mean": float(np.mean(times)), "min": float(np.min(times)), "max": float(np.max(times))},
            "path_cost_m_mean": float(np.mean(costs)) if costs else None}


def single_search_resident(gbp, t, start, goal, K, best, iters=400, seed=1, query=7):
    """ONE search resident on the device (gbp_plan_batch with one query): wall clock of the host-pointer call over the extends it ran
    (the valid random states among the STATE cells it consumed) -> microseconds per extend (+ its connect when the tree grew)."""
    P = gbp.PlanParams(K, best, iters, 4096, 0, 0, 0)
    t.plan_batch(start[None], goal[None], seed, 999, gbp.PlanParams(K, best, 3, 4096, 0, 0, 0))  # warm-up
    dt, st = None, None
    for _ in range(3):
        t0 = time.perf_counter()
        st = t.plan_batch(start[None], goal[None], seed, query, P)
        d = time.perf_counter() - t0
        dt = d if dt is None else min(dt, d)
    rs = t.sample_states(seed, query, 0, 2 * int(st["iters"][0]))
    extends = max(int(t.valid_states(rs, gbp.STANCE)[0].sum()), 1)
    return {"form": t.plan_batch_form(P, 1), "us_per_extend_and_connect": dt / extends * 1e6, "extends": extends, "iterations": int(st["iters"][0]),
            "vertices_added": int(st["nv_a"][0] + st["nv_b"][0]) - 2, "validated_actions_per_s": float(st["pair_checks"][0]) / dt,
            "solved": int(st["solved"][0]), "seconds": dt, "api": "gbp_plan_batch, one query (host pointers; includes the call's upload, launch and read-back)"}


def run_rough_k4096(gbp, torch, dev, nq=2368, iters=400, want_cpu=True):
    """BASELINE configs[1]: RRT-Connect on the reference's data/rough_terrain (committed as
    tests/golden/terrain_rough_terrain.npz), 4096 candidate actions per extend (closest valid), start (0,0) ->
    goal (8,0) at body height 0.375 m; nq independent searches (distinct Philox streams: one wave of
    the planner's resident warps), each run to its first solution or `iters` iterations."""
    t, start, goal = shipped_query(gbp, "rough_terrain", 0.375)
    s = torch.from_numpy(np.repeat(start[None], nq, 0)).to(dev); g = torch.from_numpy(np.repeat(goal[None], nq, 0)).to(dev)
    P = gbp.PlanParams(4096, 1, iters, 512, 0, 0, 0)
    dstats = torch.zeros(nq * 80, dtype=torch.uint8, device=dev)
    cur = torch.cuda.current_stream().cuda_stream
    t.plan_batch_dev(64, s.data_ptr(), g.data_ptr(), 1, 1 << 30, gbp.PlanParams(4096, 1, 4, 512, 0, 0, 0), dstats.data_ptr(), cur)  # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t.plan_batch_dev(nq, s.data_ptr(), g.data_ptr(), 1, 0, P, dstats.data_ptr(), cur)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    st = dstats.cpu().numpy().view(gbp.PLAN_STATS_DTYPE)
    ok = st["solved"] == 1
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
    out = {"workload": f"{nq} searches on data/rough_terrain, (0,0)->(8,0) at 0.375 m, K=4096 closest-valid candidates per extend, up to {iters} iterations each",
           "gpu": {"validated_actions_per_s": float(st["pair_checks"].sum() / secs), "extends_per_s": float(st["nn_queries"].sum() / 2 / secs),
                   "solved": int(ok.sum()), "searches": nq, "solved_plans_per_s": float(ok.sum() / secs), "seconds": secs,
                   "iterations_to_solve_mean": float(st["iters"][ok].mean()) if ok.any() else None,
                   "path_length_m_mean": float(st["path_length"][ok].mean()) if ok.any() else None,
                   "first_solution": first_solution_rounds(t, start, goal, gbp.PlanParams(4096, 1, iters, 512, 0, 0, 0, 1), 1184, 3, 8),
                   "single_search": single,
                   "single_search_resident": {"K4096_closest_valid": single_search_resident(gbp, t, start, goal, 4096, 1, iters),
                                              "K6_first_valid": single_search_resident(gbp, t, start, goal, 6, 0, iters)}},
           "solved_flag": bool(ok.any())}
    if want_cpu:  # the reference's own extend takes 6 candidates (NUM_GEN_STATES): its planner on the same query, all host cores
        out["reference_cpu"] = reference_cpu_planner("rough_terrain", 0, start, goal, 1.0, 25.0)
    return out


def run_slope_config0(gbp, rounds=8, want_cpu=True):
    """BASELINE configs[0]: RRT-Connect on the reference's data/slope (tests/golden/terrain_slope.npz), (0,0) -> (8,0) at body
    height 0.30 m (the fork's 0.375 m start pose is invalid on this map), default parameters
    (config/params.yaml: K = NUM_GEN_STATES = 6, first valid decides, every fork option off)."""
    t, start, goal = shipped_query(gbp, "slope", 0.30)
    nq = 2048  # the drop-in's defaults (rrt_connect.h): 2048 attempts x 32000 iterations per launch, pipelined form with 16 speculated halves per round
    S, G = np.repeat(start[None], nq, 0), np.repeat(goal[None], nq, 0)
    t.plan_batch(S[:64], G[:64], 1, 0, gbp.PlanParams(6, 0, 10, 2048, 0, 0, 0, 0))  # warm-up
    P = gbp.PlanParams(6, 0, 32000, 2048, 0, 0, 0, 1)
    t.plan_batch(S, G, 1, 1 << 40, gbp.PlanParams(6, 0, 40, 2048, 0, 0, 0, 1))  # warm-up of the form the timed launches take (its arena comes from the stream's pool)
    gpu = first_solution_rounds(t, start, goal, P, nq, rounds)
    gpu["planner_form"] = t.plan_batch_form(P, nq)
    out = {"workload": "data/slope, (0,0)->(8,0), body 0.30 m, K=6 first-valid, default params.yaml (host-pointer C ABI, wall clock)", "gpu": gpu,
           "solved_flag": gpu["calls_solved"] > 0}
    if want_cpu:
        out["reference_cpu"] = reference_cpu_planner("slope", 0, start, goal, 1.0, 25.0)
    return out


def run_rough_star_config2(gbp, budget_s=2.0, want_cpu=True):
    """BASELINE configs[2]: RRT*-Connect (choose parent + near-set rewiring, delta = 3 m, rrt_star_connect.cpp:12-75, rrt_star_connect.h:59)
    on data/rough_terrain under a fixed WALL-CLOCK planning budget: rounds of device-resident attempts at the (0,0)->(8,0) query
    (K = 64 closest-valid candidates per extend: with the reference's K = 6 its own planner completes no plan in 120 s, BASELINE.md
    row 8) until the budget is used; the cheapest post-processed path is kept, as buildRRTStarConnect keeps its best."""
    t, start, goal = shipped_query(gbp, "rough_terrain", 0.375)
    nq = 2368
    S, G = np.repeat(start[None], nq, 0), np.repeat(goal[None], nq, 0)
    P = gbp.PlanParams(64, 1, 1500, 1024, 0, 1, 1, 8)
    t.plan_batch(S[:64], G[:64], 1, 1 << 30, gbp.PlanParams(64, 1, 4, 1024, 0, 1, 1, 0))  # warm-up
    t0 = time.perf_counter()
    q0, launches, solved, best, first, checks = 0, 0, 0, None, None, 0
    while time.perf_counter() - t0 < budget_s:
        st = t.plan_batch(S, G, 1, q0, P)
        q0 += nq
        launches += 1
        ok = st["solved"] == 1
        checks += int(st["pair_checks"].sum())
        if ok.any():
            solved += int(ok.sum())
            if first is None:
                first = time.perf_counter() - t0
            c = float(st["path_cost"][ok].min())
            best = c if best is None else min(best, c)
    wall = time.perf_counter() - t0
    out = {"workload": f"data/rough_terrain, (0,0)->(8,0) at 0.375 m, RRT*-Connect, wall-clock budget {budget_s} s, K=64 closest-valid, post-processed paths",
           "gpu": {"budget_s": budget_s, "wall_s": wall, "launches": launches, "attempts_per_launch": nq, "solved_plans": solved,
                   "solved_plans_per_s": solved / wall, "time_to_first_solution_s": first, "best_path_cost_m": best,
                   "validated_actions_per_s": checks / wall},
           "solved_flag": solved > 0}
    if want_cpu:
        out["reference_cpu"] = reference_cpu_planner("rough_terrain", 1, start, goal, budget_s, 25.0)
    return out


def _pin_worker(args):
    """the unmodified reference's runRRTConnect (oracle/_ref/libgbp_ref_pin.so) on a slice of the configs[4] queries"""
    lo, hi, s, g, seed, query0, iters = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    x, y, z = rough_terrain()
    pin = po.RefPin(po.Terrain(x, y, z))
    pp = po.PinParams(0, iters, 0, 1, 0, 0, 0, 0, 0.0, 0.0, 1.0, 1.0)
    out = []
    for i in range(lo, hi):
        r, ta, tb, ps, pa = pin.run(s[i], g[i], seed, query0 + i, pp, cap=MAX_VERTS + 8)
        out.append((i, r.solved, (r.cells_used + 1) // 2, r.nv_a, r.nv_b, r.path_length))
    return out


def timed_batch(gbp, torch, dist, dev, rank, world, t, s, g, seed, query0, P):
    """one device-resident batch: every rank plans its own queries; max over ranks; NCCL gather of the statistics records"""
    nq = len(s)
    ds = torch.from_numpy(s).to(dev); dg = torch.from_numpy(g).to(dev)
    dstats = torch.zeros(nq * 80, dtype=torch.uint8, device=dev)
    cur = torch.cuda.current_stream().cuda_stream
    t.plan_batch_dev(nq, ds.data_ptr(), dg.data_ptr(), seed, query0, P, dstats.data_ptr(), cur)  # warm-up at full size (the tree arena comes from the stream-ordered pool)
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
        return None, None
    return allstats.view(gbp.PLAN_STATS_DTYPE), float(ms.item()) * 1e-3


def batch_summary(st, secs):
    ok = st["solved"] == 1
    return {"queries": int(len(st)), "solved": int(ok.sum()), "solved_plans_per_s": float(ok.sum() / secs), "queries_per_s": float(len(st) / secs),
            "validated_actions_per_s": float(st["pair_checks"].sum() / secs), "nn_queries_per_s": float(st["nn_queries"].sum() / secs), "seconds": secs,
            "mean_path_length_m": float(st["path_length"][ok].mean()) if ok.any() else None, "mean_iters": float(st["iters"].mean())}


def run(gbp, torch, dist, dev, rank, world, q_per_gpu=Q_PER_GPU, cpu_seconds=8.0, want_cpu=True):
    x, y, z = rough_terrain()
    t = gbp.Terrain(x, y, z)
    seed, stream = 11, 7000 + rank

    def valid(q):
        return t.valid_states(q, gbp.STANCE)[0]

    def queries(lo, hi, strm):
        s, g = make_queries(valid, t.sample_states, 2 * q_per_gpu, seed, strm, lo, hi)
        hs, _ = t.ground_height(s[:, 0], s[:, 1]); hg, _ = t.ground_height(g[:, 0], g[:, 1])
        g[:, 2] = s[:, 2] - hs + hg  # goal height: body height of the start above the goal's own ground
        vg = valid(g)
        return s[vg == 1][:q_per_gpu], g[vg == 1][:q_per_gpu]

    s, g = queries(4.0, 8.0, stream)
    nq = len(s)
    P = gbp.PlanParams(K_CAND, 0, MAX_ITERS, MAX_VERTS, 0, 0, 0)
    query0 = rank * q_per_gpu
    st, secs = timed_batch(gbp, torch, dist, dev, rank, world, t, s, g, seed, query0, P)
    s35, g35 = queries(3.0, 5.0, stream + 500)
    st35, secs35 = timed_batch(gbp, torch, dist, dev, rank, world, t, s35, g35, seed, query0 + (1 << 32), P)
    if rank != 0:
        return None
    out = {"workload": f"{nq} RRT-Connect queries per GPU (BASELINE configs[4]), synthetic rough terrain {MAP_N}x{MAP_N} @ {PITCH} m, "
                       f"blocks {BLOCK * PITCH:.1f} m, steps U(0,{STEP_H}) m; start/goal 4-8 m apart (SURVEY 8d config 5); K={K_CAND} first-valid, "
                       f"budget {MAX_ITERS} iterations / {MAX_VERTS} vertices per tree"}
    out.update(batch_summary(st, secs))
    out["planner_form"] = t.plan_batch_form(P, nq)
    out["stats_gather_bytes"] = int(st.nbytes)
    out["separation_3_5m"] = dict(batch_summary(st35, secs35), note="round 1's workload (start/goal 3-5 m apart), same budget, for comparison")
    cfg = {}
    cfg["0_slope_rrt_connect_default_params"] = run_slope_config0(gbp, want_cpu=want_cpu)
    cfg["1_rough_terrain_k4096"] = run_rough_k4096(gbp, torch, dev, want_cpu=want_cpu)
    cfg["2_rough_terrain_rrt_star_time_budget"] = run_rough_star_config2(gbp, want_cpu=want_cpu)
    out["configs"] = cfg
    # RRT*-Connect against RRT-Connect at the same iteration budget on the first queries of the configs[4] set (post-processed paths)
    ns, star_iters = min(nq, 2048), 400
    res = {}
    cur = torch.cuda.current_stream().cuda_stream
    ds = torch.from_numpy(s[:ns]).to(dev); dg = torch.from_numpy(g[:ns]).to(dev)
    dstats = torch.zeros(ns * 80, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, PP in (("rrt_star_connect", gbp.PlanParams(K_CAND, 0, star_iters, MAX_VERTS, 0, 1, 1)),
                     ("rrt_connect_same_budget", gbp.PlanParams(K_CAND, 0, star_iters, MAX_VERTS, 0, 0, 1))):
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
        cores = os.cpu_count() or 1
        # (a) the UNMODIFIED reference's loops on the same queries and the same Philox stream (pin library: samplers served from the
        #     stream, iteration budget through the sampler), one process per core, on a bounded sample
        if po.RefPin.available():
            import multiprocessing as mp
            procs = min(cores, 16)
            m = 2 * procs
            t0 = time.perf_counter()
            with mp.get_context("spawn").Pool(procs) as pool:
                rows = sum(pool.map(_pin_worker, [(k * m // procs, (k + 1) * m // procs, s, g, seed, 0, MAX_ITERS) for k in range(procs)]), [])
            dt = time.perf_counter() - t0
            rows.sort()
            gst = st[:m]
            same = all((r[1], r[2], r[3], r[4]) == (int(gst["solved"][r[0]]), int(gst["iters"][r[0]]), int(gst["nv_a"][r[0]]), int(gst["nv_b"][r[0]])) and
                       (not r[1] or np.float64(r[5]).view(np.uint64) == np.float64(gst["path_length"][r[0]]).view(np.uint64)) for r in rows)
            nsolved = sum(r[1] for r in rows)
            out["reference_cpu"] = {"kind": "reference", "cores": procs, "queries": m, "solved": int(nsolved), "solved_plans_per_s": nsolved / dt,
                                    "queries_per_s": m / dt, "seconds": dt,
                                    "note": "unmodified runRRTConnect / extend / newConfig / connect (oracle/_ref/libgbp_ref_pin.so) on the first queries of "
                                            "rank 0, same Philox stream and iteration budget; includes process start-up",
                                    "gpu_trees_identical_on_sample": bool(same)}
        # (b) the oracle restatement (port) on all cores: the faster CPU baseline
        o = po.Oracle(po.Terrain(x, y, z))
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
                               "note": "iteration-budgeted oracle planner (pinned to the reference's loops, tests/test_oracle_vs_ref.py), same Philox stream",
                               "gpu_stats_identical_on_sample": bool(same),
                               "path_length_max_rel_diff": float(np.max(np.abs(cst["path_length"] - gst["path_length"]) /
                                                                        np.maximum(1e-300, np.abs(cst["path_length"]))) if m else 0.0)}
    return out
