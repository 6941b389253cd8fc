#!/bin/bash
# A/B of the pipelined planner's knobs on configs[4]: GBP_PIPE_SPEC (half-iterations a query speculates per round) and
# GBP_PIPE_RESUME (running queries at which the rest of the batch moves to k_pipe_resume; 0 = never)
mkdir -p gpurun_out
L=gpurun_out/spec_sweep.log
: > $L
run() { echo "== $*" >> $L; env "${@:3}" GBP_PIPE_TRACE=1 timeout 300 python tools/bench_planner_modes.py $1 $2 >> $L 2>&1; }
run 65536 pipe GBP_PIPE_RESUME=0
run 65536 pipe GBP_PIPE_RESUME=2048
run 65536 pipe GBP_X=0
run 65536 pipe GBP_PIPE_RESUME=8192
run 65536 pipe GBP_PIPE_RESUME=16384
run 16384 pipe GBP_X=0
run 8192 pipe GBP_X=0
