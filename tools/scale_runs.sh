#!/bin/bash
# weak (ER500) and strong (ER10000, 128 trajectories split over the ranks) scaling lines at N GPUs; run under gpurun --gpus N
N=$1; TAG=${2:-r03}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$R --master-port 29701 bench.py --gpus $N --steps 400 --warmup 16 > gpurun_out/${TAG}_weak_er500_n$N.log 2>gpurun_out/${TAG}_weak_er500_n$N.err
$R --master-port 29702 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_weak_er500_n${N}_k20.log 2>/dev/null
$R --master-port 29703 bench.py --gpus $N --steps 400 --warmup 16 --workload er10000 --scaling strong > gpurun_out/${TAG}_strong_er10000_n$N.log 2>gpurun_out/${TAG}_strong_er10000_n$N.err
$R --master-port 29704 bench.py --gpus $N --steps 400 --warmup 16 --workload er2000 --scaling strong > gpurun_out/${TAG}_strong_er2000_n$N.log 2>/dev/null
echo done $N
