#!/bin/bash
# ncu evidence for the pipelined planner at 65,536 queries (configs[4] queries): a launch list of ~6 rounds and one
# --set full capture each of k_walk_seg and k_pipe_prep around round 100.  Run only after the same command exited 0 without ncu.
mkdir -p gpurun_out
CMD="python tools/bench_planner_modes.py 65536 pipe"
$CMD > gpurun_out/pipe_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 36 --csv --log-file gpurun_out/r2b_pipe_launches.csv $CMD > gpurun_out/ncu_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_walk_seg -s 100 -c 1 -f -o gpurun_out/pipe_walk $CMD > gpurun_out/ncu_p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pipe_prep -s 100 -c 1 -f -o gpurun_out/pipe_prep $CMD > gpurun_out/ncu_p3.log 2>&1
python tools/ncu_summary.py gpurun_out/pipe_walk.ncu-rep > gpurun_out/r2b_walk_seg_ncu_summary.csv
python tools/ncu_summary.py gpurun_out/pipe_walk.ncu-rep --lines k_walk_seg --top 40 > gpurun_out/r2b_walk_seg_lines.txt
python tools/ncu_summary.py gpurun_out/pipe_prep.ncu-rep > gpurun_out/r2b_pipe_prep_ncu_summary.csv
python tools/ncu_summary.py gpurun_out/pipe_prep.ncu-rep --lines k_pipe_prep --top 40 > gpurun_out/r2b_pipe_prep_lines.txt
rm -f gpurun_out/pipe_walk.ncu-rep gpurun_out/pipe_prep.ncu-rep  # (gpurun brings back at most 64 MiB)
