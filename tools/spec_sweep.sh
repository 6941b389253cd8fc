#!/bin/bash
# A/B of the pipelined planner's knobs on configs[4]
mkdir -p gpurun_out
L=gpurun_out/spec_sweep.log
: > $L
run() { echo "== $*" >> $L; env "${@:3}" GBP_PIPE_TRACE=1 timeout 300 python tools/bench_planner_modes.py $1 $2 >> $L 2>&1; }
run 65536 pipe GBP_PIPE_SIDE_CTAS=4
run 65536 pipe GBP_X=1
run 65536 pipe GBP_PIPE_SIDE_CTAS=4
run 65536 pipe GBP_X=1
run 16384 pipe GBP_PIPE_SIDE_CTAS=4
run 16384 pipe GBP_X=1
