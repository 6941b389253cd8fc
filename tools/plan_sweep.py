"""Solved plans/s of the device-resident planner against the number of queries per launch (tail / wave effects)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench_plans  # noqa: E402
import global_body_planner_b200 as gbp  # noqa: E402

dev = torch.device("cuda", 0)
bench_plans.run_rough_k4096 = lambda *a, **k: None
bench_plans.run_slope_config0 = lambda *a, **k: None
for q in [int(a) for a in sys.argv[1:]] or [8192, 65536]:
    r = bench_plans.run(gbp, torch, None, dev, 0, 1, q_per_gpu=q, want_cpu=False)
    print(json.dumps({k: r[k] for k in ("queries", "solved", "solved_plans_per_s", "queries_per_s", "seconds", "mean_iters")}))
