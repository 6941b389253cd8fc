"""The N>1 launch logic of bench.py on CPU: world_size-2 gloo process group, contiguous query /
candidate ranges per rank, disjoint Philox streams, final gather of fixed-size statistics records.
(The GPU work itself is per-rank and collective-free)"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q_total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = po.Terrain.from_npz(os.path.join(ROOT, "tests", "golden", "terrain_slope.npz"))
    o = po.Oracle(T)
    per = q_total // world
    lo = rank * per  # contiguous block of queries per rank, rank = GPU index
    # queries are a pure function of the global query id, so any sharding yields the same records
    q = o.sample_states(3, 1, 0, 4000)
    q[:, 3:8] = 0; q[:, 3] = 0.5
    v, _ = o.valid_states(q, po.STANCE)
    q = q[v == 1]
    starts, goals = q[lo:lo + per], q[lo + 40:lo + 40 + per]
    P = po.PlanParams(6, 0, 60, 64, 0, 0, 0)
    st = o.plan_batch(starts, goals, 7, lo, P, 1)  # query id = global index (the Philox stream)
    rec = torch.from_numpy(st.view(np.uint8).copy())
    gathered = [torch.empty_like(rec) for _ in range(world)] if rank == 0 else None
    dist.gather(rec, gathered, dst=0)  # the final statistics gather (NCCL on the GPU box)
    tmax = torch.tensor([float(rank + 1)])
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)  # max-over-ranks timing
    if rank == 0:
        allst = torch.cat(gathered).numpy().view(po.PLAN_STATS_DTYPE)
        np.save(out_path, allst)
        assert tmax.item() == world
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_plan_statistics_equal_single_rank(tmp_path):
    out = str(tmp_path / "stats.npy")
    q_total = 16
    mp.spawn(_worker, args=(2, _free_port(), q_total, out), nprocs=2, join=True)
    sharded = np.load(out)
    T = po.Terrain.from_npz(os.path.join(ROOT, "tests", "golden", "terrain_slope.npz"))
    o = po.Oracle(T)
    q = o.sample_states(3, 1, 0, 4000)
    q[:, 3:8] = 0; q[:, 3] = 0.5
    v, _ = o.valid_states(q, po.STANCE)
    q = q[v == 1]
    P = po.PlanParams(6, 0, 60, 64, 0, 0, 0)
    single = o.plan_batch(q[:q_total], q[40:40 + q_total], 7, 0, P, 1)
    assert len(sharded) == q_total
    for key in single.dtype.names:
        assert np.array_equal(sharded[key], single[key]), key


def test_bench_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the reference arm; the others exit 0 without output."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
