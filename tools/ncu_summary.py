#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the CSV kept under profiles/ (kernel, metric, unit, value), plus the
per-source-line instruction / stall-sample breakdown of one kernel (needs -lineinfo and --import-source on).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>_ncu_summary.csv
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep --lines k_walk_mixed --top 40 > profiles/<name>_lines.txt
"""
import argparse
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_sectors_mem_texture.sum",
        "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def summary(rep):
    rows = list(csv.reader(ncu("-i", rep, "--page", "raw", "--csv").splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    keys = KEYS + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "metric", "unit", "value"])
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("gbp::", "")
        for k in keys:
            if k in ix:
                w.writerow([name, k, units[ix[k]], r[ix[k]]])


def lines(rep, kernel, top):
    out = ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kernel)
    cur, agg = None, {}
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        try:
            ln, smp, inst, thr = int(r[0]), int(r[6]), int(r[7]), int(r[8])
        except (ValueError, IndexError):
            continue
        a = agg.setdefault((cur, ln), [0, 0, 0, r[1].strip()])
        a[0] += inst; a[1] += smp; a[2] += thr
    tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
    print(f"# {kernel}: warp instructions {tot}, stall samples {ts}; per source line: share of instructions, active threads, share of samples")
    files = {}
    for k, a in agg.items():
        files[k[0]] = files.get(k[0], 0) + a[0]
    print("# by file:", {k: f"{v / tot:.1%}" for k, v in sorted(files.items(), key=lambda kv: -kv[1])[:4]})
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{k[0]:16s}:{k[1]:4d} inst {a[0] / tot:6.2%} thr {a[2] / max(a[0], 1):5.1f} smp {a[1] / max(ts, 1):6.2%}  {a[3][:100]}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--lines", default=None, help="kernel-name regex: print the per-source-line breakdown instead of the summary")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    lines(a.report, a.lines, a.top) if a.lines else summary(a.report)
