#!/usr/bin/env python
"""Memcpy-only probe of the host side of the end-to-end path: every rank copies pinned host buffers to / from its GPU
at the same time (torchrun, one rank per GPU) and rank 0 prints per-rank and aggregate GB/s for H2D alone, D2H alone and
both at once — the ceiling any host-pointer call of N ranks on this box shares.  Also prints `nvidia-smi topo -m`.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/probe_h2d.py
"""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 256 << 20
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev); d_b = torch.ones(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for rep in range(2):  # first pass warms up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(8):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_a.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        per_dir = 8 * nbytes / dt / 1e9
        res[mode] = per_dir
    t = torch.tensor([res["h2d"], res["d2h"], res["both"]], dtype=torch.float64, device=dev)
    if world > 1:
        g = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, g, dst=0)
        allr = torch.stack(g).cpu().numpy() if rank == 0 else None
    else:
        allr = t.cpu().numpy()[None]
    if rank == 0:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        out = {"n_gpus": world, "bytes_per_copy": nbytes,
               "h2d_GBps_per_rank": [round(float(v), 1) for v in allr[:, 0]], "h2d_GBps_aggregate": float(allr[:, 0].sum()),
               "d2h_GBps_per_rank": [round(float(v), 1) for v in allr[:, 1]], "d2h_GBps_aggregate": float(allr[:, 1].sum()),
               "both_GBps_per_direction_per_rank": [round(float(v), 1) for v in allr[:, 2]], "both_GBps_per_direction_aggregate": float(allr[:, 2].sum()),
               "host_cpus": os.cpu_count(), "topology": topo}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
