#!/bin/bash
# BASELINE configs[0] (data/slope): time to the first solution, megakernel against the pipelined form with speculation
mkdir -p gpurun_out
L=gpurun_out/ttfs_pipe.log
: > $L
run() { echo "== $*" >> $L; env "${@:2}" timeout 300 python tools/ttfs_sweep.py 8 $1 >> $L 2>&1; }
run "3552:8000:2048" GBP_PLAN_MODE=mega
run "2048:8000:2048 4096:8000:2048 8192:8000:2048 4096:16000:2048 2048:32000:2048 4096:32000:2048 1024:64000:2048" GBP_PLAN_MODE=pipe
