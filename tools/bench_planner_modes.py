#!/usr/bin/env python
"""A/B of the forms of the batch planner (GBP_PLAN_MODE=mega | pipe | step) on bench_plans' configs[4] queries: identical
statistics required, plans/s of each.  usage: python tools/bench_planner_modes.py [queries [modes [query stream]]]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench_plans as bp
    import global_body_planner_b200 as gbp
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    dev = torch.device("cuda", 0)
    x, y, z = bp.rough_terrain()
    t = gbp.Terrain(x, y, z)
    valid = lambda q: t.valid_states(q, gbp.STANCE)[0]
    stream = int(sys.argv[3]) if len(sys.argv) > 3 else 7000  # bench.py: 7000 + rank
    s, g = bp.make_queries(valid, t.sample_states, 2 * nq, 11, stream, 4.0, 8.0)
    hs, _ = t.ground_height(s[:, 0], s[:, 1]); hg, _ = t.ground_height(g[:, 0], g[:, 1])
    g[:, 2] = s[:, 2] - hs + hg
    vg = valid(g)
    s, g = s[vg == 1][:nq], g[vg == 1][:nq]
    P = gbp.PlanParams(bp.K_CAND, 0, bp.MAX_ITERS, bp.MAX_VERTS, 0, 0, 0)
    out, ref = {}, None
    for mode in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("mega", "pipe")):
        os.environ["GBP_PLAN_MODE"] = mode
        st, secs = bp.timed_batch(gbp, torch, None, dev, 0, 1, t, s, g, 11, 0, P)
        for _ in range(int(os.environ.get("GBP_BENCH_REPEAT", "1")) - 1):  # run-to-run spread of the same batch
            secs2 = bp.timed_batch(gbp, torch, None, dev, 0, 1, t, s, g, 11, 0, P)[1]
            print(f"{mode}: repeat {secs2:.4f} s", file=sys.stderr)
        out[mode] = bp.batch_summary(st, secs)
        if ref is None:
            ref = st.copy()
        else:
            out["identical"] = all(np.array_equal(ref[k], st[k]) for k in ref.dtype.names if k not in ("path_yaw", "path_cost")) and \
                bool(np.allclose(ref["path_yaw"], st["path_yaw"], rtol=0, atol=1e-9))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
