#!/usr/bin/env python
"""Times the narrow-wire sample + validate call on bench.py's configs[3] batch: resident (gbp_sample_validate_dev, CUDA
events) and end to end (gbp_sample_validate, pinned host buffers), and cross-checks its verdict bits against the dense
call's verdict bytes on the same candidates.  One JSON line."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=16_777_216)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-idx", action="store_true", help="implicit rows (state_idx = NULL)")
    ap.add_argument("--skip-dense", action="store_true")
    ap.add_argument("--no-promise", action="store_true", help="start_states_valid = 0")
    args = ap.parse_args()
    import torch
    import bench
    import global_body_planner_b200 as gbp
    from global_body_planner_b200 import capi
    dev = torch.device("cuda", 0)
    x, y, z = bench.synthetic_map()
    t = gbp.Terrain(x, y, z)
    n = args.candidates
    states, actions, direction = bench.device_batch(torch, capi, t, x, y, n, 1, 100, dev)
    cur = torch.cuda.current_stream().cuda_stream
    out = {"n": n}
    if not args.skip_dense:
        verdict = torch.empty(n, dtype=torch.uint8, device=dev)
        t.validate_pairs_dev(n, states.data_ptr(), actions.data_ptr(), direction.data_ptr(), 0, 0, verdict.data_ptr(), 0, 0, 0, cur)
        torch.cuda.synchronize()
        dense_cnt = t.validate_counters()
    del actions
    idx = torch.arange(n, dtype=torch.int32, device=dev)
    nw = (n + 31) // 32
    bits = torch.empty(nw, dtype=torch.int32, device=dev)
    cap = n // 64
    vi = torch.empty(cap, dtype=torch.int32, device=dev); sn = torch.empty((cap, 8), dtype=torch.float64, device=dev)
    tn = torch.empty(cap, dtype=torch.float64, device=dev); ac = torch.empty((cap, 10), dtype=torch.float64, device=dev)
    res = torch.zeros(8, dtype=torch.int64, device=dev)
    p = gbp.sv_params(1, 101, 0, states_valid="--no-promise" not in sys.argv)  # bench.device_batch samples the actions from (seed 1, stream 101, idx 0 ...)

    def step():
        t.sample_validate_dev(states.data_ptr(), n, n, p, 0 if args.no_idx else idx.data_ptr(), direction.data_ptr(), bits.data_ptr(), 0, cap,
                              vi.data_ptr(), sn.data_ptr(), tn.data_ptr(), ac.data_ptr(), res.data_ptr(), cur)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    r = res.cpu().numpy()
    out.update(resident_ms=float(np.mean(ms)), resident_ms_min=float(np.min(ms)), resident_G_per_s=n / np.mean(ms) / 1e6,
               n_valid=int(r[0]), k_mean=r[1] / n, L_mean=r[2] / n, oog=int(r[4]), near=int(r[5]))
    if not args.skip_dense:
        got = torch.from_numpy(gbp.unpack_bits(bits.cpu().numpy().view(np.uint32), n)).to(dev)
        out["bits_equal_dense_verdicts"] = bool(torch.equal(got, verdict))
        out["counters_equal_dense"] = (int(r[1]), int(r[2]), int(r[3])) == (dense_cnt["substates"], dense_cnt["lookups"], dense_cnt["nanprobes"])
        del verdict, got
    # ---- end to end: host pointers, pinned
    tab = gbp.States.__new__(gbp.States)  # table uploaded once (device-to-host-to-device round trip avoided: reuse the rows)
    hstates = states.cpu().numpy()
    tab = gbp.States(hstates)
    del hstates
    hidx = torch.empty(n, dtype=torch.int32, pin_memory=True); hidx.copy_(idx)
    hdir = torch.empty(n, dtype=torch.uint8, pin_memory=True); hdir.copy_(direction)
    hbits = torch.empty(nw, dtype=torch.int32, pin_memory=True)
    hvi = torch.empty(cap, dtype=torch.int32, pin_memory=True); hsn = torch.empty((cap, 8), dtype=torch.float64, pin_memory=True)
    htn = torch.empty(cap, dtype=torch.float64, pin_memory=True); hac = torch.empty((cap, 10), dtype=torch.float64, pin_memory=True)
    hres = gbp.SvResult()
    L = gbp.lib()
    vp = lambda tt: C.c_void_p(tt.data_ptr())

    def e2e():
        rc = L.gbp_sample_validate(t.h, tab.h, C.c_int64(n), None if args.no_idx else vp(hidx), vp(hdir), C.byref(p), vp(hbits), None,
                                   C.c_int64(cap), vp(hvi), vp(hsn), vp(htn), vp(hac), C.byref(hres))
        assert rc == 0, L.gbp_last_error()

    e2e()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e()
    dt = (time.perf_counter() - t0) / args.steps
    out.update(e2e_ms=dt * 1e3, e2e_G_per_s=n / dt / 1e9, e2e_n_valid=int(hres.n_valid),
               e2e_bits_equal_resident=bool(torch.equal(hbits, bits.cpu())),
               e2e_rows_equal_resident=bool(torch.equal(hvi[:hres.n_valid], vi[:hres.n_valid].cpu()) and
                                            torch.equal(hsn[:hres.n_valid].view(torch.int64), sn[:hres.n_valid].cpu().view(torch.int64))),
               h2d_bytes=n * ((0 if args.no_idx else 4) + 1), d2h_bytes=nw * 4 + int(hres.n_valid) * (4 + 64 + 8 + 80) + 64)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
