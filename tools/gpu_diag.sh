#!/bin/bash
# Run every diagnostic section in its own process under a timeout (a hung kernel must not eat the GPU lease).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/diag_smi.log 2>&1
for s in ${@:-env model tc rollout perf}; do
  echo "##### section $s" | tee -a gpurun_out/diag.log
  timeout 300 python tools/gpu_diag.py $s >> gpurun_out/diag.log 2>&1
  echo "##### section $s exit $?" | tee -a gpurun_out/diag.log
done
tail -c 3000 gpurun_out/diag.log
