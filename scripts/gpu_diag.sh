#!/bin/bash
# run every diagnostic group in its own process, each under a timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/diag.log
for g in "$@"; do
  timeout 600 python scripts/gpu_diag.py $g 2>&1 | tail -40 | tee -a gpurun_out/diag.log
done
