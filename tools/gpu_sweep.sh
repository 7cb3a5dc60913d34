#!/bin/bash
# Parity first, then a (scenes-per-CTA, threads) sweep of the step kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for cfg in "4 256" "4 192" "4 160" "3 128" "2 96" "2 128" "6 256" "5 224" "1 64"; do
  set -- $cfg
  echo "== group $1 threads $2"
  B2C_ENV_GROUP=$1 B2C_ENV_THREADS=$2 timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
done | tee gpurun_out/sweep.log
B2C_ENV_THREADS=256 timeout 120 python tools/env_perf.py 4096 10 parking_lot 2>&1 | tail -1 | tee -a gpurun_out/sweep.log
timeout 120 python tools/env_perf.py 4096 40 tollgate 2>&1 | tail -1 | tee -a gpurun_out/sweep.log
timeout 120 python tools/env_perf.py 4096 40 roundabout 2>&1 | tail -1 | tee -a gpurun_out/sweep.log
