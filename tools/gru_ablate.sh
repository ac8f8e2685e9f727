# split GRU kernel alone: tile width of the leftover tiles (auto / 64 / 32 / 16 units) at three batch sizes
for wl in er500 er10000; do for n in "" 4 2 1; do ECORD_GRU_NARROW=$n python tools/gru_time.py $wl bf16x3 1 2>&1 | tail -1; done; done
