# ring-depth sensitivity of the split GRU kernel: TMA only (131), TMA + MMA (130), everything (0)
for v in s3 s4 s6; do for d in 131 130 0; do ECORD_B200_LIB=ecord_b200/_build/lib_$v.so ECORD_TC_DEBUG=$d python tools/gru_time.py er500 bf16x3 1 2>&1 | tail -1 | sed "s/^/$v /"; done; done
for d in 131 130 0; do ECORD_TC_DEBUG=$d python tools/gru_time.py er500 bf16x3 1 2>&1 | tail -1 | sed "s/^/s5 /"; done
