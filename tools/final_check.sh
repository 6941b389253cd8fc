#!/bin/bash
# round-end check on one GPU: tests, smoke, the default bench of both arms, then the ncu evidence of the headline step
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log | cut -c1-300
S=$(date +%s); python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
S=$(date +%s); python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "reference arm rc=$? wall=$(( $(date +%s) - S ))s"; cut -c1-400 gpurun_out/bench_final_ref.json
CMD="python bench.py --no-cpu --no-plans --steps 2 --warmup 1 --e2e-steps 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv $CMD > gpurun_out/ncu_f1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_walk_sv -s 2 -c 1 -f -o gpurun_out/final_walk $CMD > gpurun_out/ncu_f2.log 2>&1; echo "set full rc=$?"
python tools/ncu_summary.py gpurun_out/final_walk.ncu-rep > gpurun_out/r2b_walk_sv_ncu_summary.csv
python tools/ncu_summary.py gpurun_out/final_walk.ncu-rep --lines k_walk_sv --top 40 > gpurun_out/r2b_walk_sv_lines.txt
rm -f gpurun_out/final_walk.ncu-rep
grep -E "time_duration|dram__bytes|inst_executed.sum|thread_inst_executed_per" gpurun_out/r2b_walk_sv_ncu_summary.csv
