"""Time to the first solution of BASELINE configs[0] (data/slope, (0,0)->(8,0) at 0.30 m, K = 6 first-valid) as a function of
how a round of device-resident attempts is sized: attempts per launch x iteration budget x vertices per tree.
  python tools/ttfs_sweep.py [calls] [attempts:iterations:vertices ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import global_body_planner_b200 as gbp  # noqa: E402
import bench_plans as bp  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 6
combos = [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:]] or [(3552, 8000, 2048), (3552, 32000, 4096), (3552, 200000, 8192),
                                                                          (888, 200000, 8192), (14208, 32000, 4096)]
t, start, goal = bp.shipped_query(gbp, "slope", 0.30)
t.plan_batch(np.repeat(start[None], 64, 0), np.repeat(goal[None], 64, 0), 1, 0, gbp.PlanParams(6, 0, 10, 2048, 0, 0, 0, 0))
for nq, iters, cap in combos:
    S, G = np.repeat(start[None], nq, 0), np.repeat(goal[None], nq, 0)
    P = gbp.PlanParams(6, 0, iters, cap, 0, 0, 0, 1)
    times, its, nvs, launches, q0 = [], [], [], 0, nq
    for c in range(calls):
        t0 = time.perf_counter()
        for _ in range(40):
            st = t.plan_batch(S, G, 1 + c, q0, P)
            q0 += nq; launches += 1
            ok = st["solved"] == 1
            if ok.any():
                i = int(np.nonzero(ok)[0][0])
                its.append(int(st["iters"][i])); nvs.append(int(st["nv_a"][i] + st["nv_b"][i]))
                break
        times.append(time.perf_counter() - t0)
    print(f"attempts {nq} iterations {iters} vertices {cap}: first solution mean {np.mean(times):.3f} s (min {np.min(times):.3f}, max {np.max(times):.3f}), "
          f"{launches} launches in {calls} calls; solving attempt: iterations {its}, vertices {nvs}", flush=True)
