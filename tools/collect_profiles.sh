#!/bin/bash
# Round-end evidence (run under gpurun, 1 GPU): plain runs first, then the ncu launch lists and full captures.
set -u
mkdir -p gpurun_out
O=gpurun_out
T="timeout -s KILL"
$T 300 python bench.py --steps 64 --warmup 16 --no-cpu-baseline > $O/plain_er500.log 2>&1 &&
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_er500.csv \
    python bench.py --steps 64 --warmup 16 --no-cpu-baseline > $O/ncu_er500.log 2>&1
$T 300 python bench.py --workload er10000 --steps 32 --warmup 16 --no-cpu-baseline > $O/plain_er10000.log 2>&1 &&
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_er10000.csv \
    python bench.py --workload er10000 --steps 32 --warmup 16 --no-cpu-baseline > $O/ncu_er10000.log 2>&1
for w in er500 er10000; do
  $T 120 python tools/q_prof.py $w tc 5 > $O/qprof_$w.log 2>&1 &&
  $T 300 ncu --set full --clock-control none --import-source on -k regex:k_q_tc -s 9 -c 1 -o $O/prof_qtc_$w \
      python tools/q_prof.py $w tc 5 > $O/qprof_ncu_$w.log 2>&1
done
$T 120 python tools/p_prof.py er500 5 > $O/pprof.log 2>&1 &&
$T 300 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_tail_tc" -s 24 -c 3 -o $O/prof_policy_er500 \
    python tools/p_prof.py er500 5 > $O/pprof_ncu.log 2>&1
tail -1 $O/qprof_er500.log $O/qprof_er10000.log $O/pprof.log
