#!/bin/bash
# Round-end evidence (run under gpurun, 1 GPU): plain runs first, then the ncu launch lists and full captures.
#   bash tools/collect_profiles.sh [tag]      -> gpurun_out/<tag>_*; summarise here with tools/ncu_table.py / ncu_metrics.py
set -u
TAG=${1:-r03}
mkdir -p gpurun_out
O=gpurun_out
T="timeout -s KILL"
for w in er500 er10000; do
  $T 300 python bench.py --workload $w --steps 32 --warmup 16 --no-cpu-baseline > $O/${TAG}_plain_$w.log 2>&1 &&
  $T 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_launches_$w.csv \
      python bench.py --workload $w --steps 32 --warmup 16 --no-cpu-baseline > $O/${TAG}_ncu_$w.log 2>&1
done
# policy kernels (GRU, projection, tail), isolated launches
$T 120 python tools/p_prof.py er500 5 > $O/${TAG}_pprof.log 2>&1 &&
$T 400 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc2|k_tail_tc" -s 24 -c 3 -o $O/${TAG}_prof_policy_er500 -f \
    python tools/p_prof.py er500 5 > $O/${TAG}_pprof_ncu.log 2>&1
# Q kernel
for w in er500 er10000; do
  $T 120 python tools/q_prof.py $w tc 5 > $O/${TAG}_qprof_$w.log 2>&1 &&
  $T 400 ncu --set full --clock-control none --import-source on -k regex:k_q_tc -s 9 -c 1 -o $O/${TAG}_prof_qtc_$w -f \
      python tools/q_prof.py $w tc 5 > $O/${TAG}_qprof_ncu_$w.log 2>&1
done
# HBM-bound kernels and the t=0 set-up (GNN, env init, env step, selection), out of a short bench run
for w in er500 er10000; do
  $T 500 ncu --set full --clock-control none --import-source on \
      -k regex:"k_gnn_update|k_gnn_ln_send|k_gnn_finish|k_env_init_peek|k_env_init_score_peek|k_env_step_fused|k_select_finish" -c 14 \
      -o $O/${TAG}_prof_misc_$w -f python bench.py --workload $w --steps 16 --warmup 3 --no-cpu-baseline > $O/${TAG}_misc_ncu_$w.log 2>&1
done
tail -1 $O/${TAG}_pprof.log $O/${TAG}_qprof_er500.log $O/${TAG}_qprof_er10000.log
