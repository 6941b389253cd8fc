"""Latency of ONE RRT-Connect search driven extend by extend through the host-pointer call gbp_extend on data/rough_terrain
(BASELINE configs[1]: K = 4096 candidate actions per extend).  `ncu --metrics gpu__time_duration.sum` on this script
gives the kernel split of one extend."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import global_body_planner_b200 as gbp  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_ext = int(sys.argv[2]) if len(sys.argv) > 2 else 600
d = np.load(os.path.join(ROOT, "tests", "golden", "terrain_rough_terrain.npz"))
t = gbp.Terrain(d["x"], d["y"], d["z"], d["dx"], d["dy"], d["dz"])
h, _ = t.ground_height([0.0], [0.0])
start = np.array([0, 0, h[0] + 0.375, 1, 0, 0, 0, 0.0])
tree = gbp.Tree(8192, start)
targets = t.sample_states(5, 77, 0, 8 * n_ext + 200)
targets = targets[t.valid_states(targets, gbp.STANCE)[0] == 1][:n_ext + 20]
for i in range(20):
    tree.extend(t, targets[i], gbp.FORWARD, K, 1, 5, 78, i * K)
t0 = time.perf_counter()
added = 0
for i in range(20, len(targets)):
    st, nid, chk = tree.extend(t, targets[i], gbp.FORWARD, K, 1, 5, 78, i * K)
    added += int(st != gbp.TRAPPED)
dt = time.perf_counter() - t0
n = len(targets) - 20
print(f"K={K}: {dt / n * 1e6:.1f} us per extend, {n * K / dt / 1e6:.1f} M validated actions/s, {added} of {n} extends added a vertex")
# gbp_connect against the grown tree (rrt_connect.cpp:98-120): nearest neighbour + attemptConnect + append per call
ctargets = targets[: min(300, len(targets))]
n0 = tree.size() if hasattr(tree, "size") else None
for tg in ctargets[:10]:
    tree.connect(t, tg, gbp.FORWARD)
t0 = time.perf_counter()
res = [tree.connect(t, tg, gbp.FORWARD)[0] for tg in ctargets]
dt = time.perf_counter() - t0
print(f"connect: {dt / len(ctargets) * 1e6:.1f} us per call, statuses trapped/advanced/reached = "
      f"{res.count(gbp.TRAPPED)}/{res.count(gbp.ADVANCED)}/{res.count(gbp.REACHED)}")
