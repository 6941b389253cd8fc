"""Latency of ONE RRT-Connect search resident on the device (gbp_plan_batch with one query) on data/rough_terrain,
(0,0)->(8,0) at 0.375 m: microseconds per extend (+ its connect when the tree grew) in the device-wide form (k_plan_wide)
and in the megakernel (one warp), at K = 6 first-valid and K = 4096 closest-valid.
  python tools/bench_wide.py [iterations]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import global_body_planner_b200 as gbp  # noqa: E402
import bench_plans as bp  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 400
t, start, goal = bp.shipped_query(gbp, "rough_terrain", 0.375)
out = {}
for K, best in ((6, 0), (64, 1), (4096, 1)):
    for mode in ("wide", "mega"):
        os.environ["GBP_PLAN_MODE"] = mode
        n_it = iters if (mode == "wide" or K <= 64) else max(20, iters // 10)
        P = gbp.PlanParams(K, best, n_it, 4096, 0, 0, 0)
        t.plan_batch(start[None], goal[None], 1, 999, gbp.PlanParams(K, best, 3, 4096, 0, 0, 0))  # warm-up
        best_dt, st = None, None
        for rep in range(3):
            t0 = time.perf_counter()
            st = t.plan_batch(start[None], goal[None], 1, 7, P)
            dt = time.perf_counter() - t0
            best_dt = dt if best_dt is None else min(best_dt, dt)
        # extends = valid random states among the STATE cells the search consumed (2 per iteration)
        cells = 2 * int(st["iters"][0])
        rs = t.sample_states(1, 7, 0, cells)
        extends = int(t.valid_states(rs, gbp.STANCE)[0].sum())
        grown = int(st["nv_a"][0] + st["nv_b"][0]) - 2
        print(f"K={K} {t.plan_batch_form(P, 1)}: {best_dt * 1e3:.2f} ms for {int(st['iters'][0])} iterations, {extends} extends (<= {grown} vertices added), "
              f"{best_dt / max(extends, 1) * 1e6:.1f} us per extend(+connect), {int(st['pair_checks'][0]) / best_dt / 1e6:.1f} M validated actions/s, solved {int(st['solved'][0])}",
              flush=True)
os.environ.pop("GBP_PLAN_MODE", None)
