#!/bin/bash
# First visit of the next round (about 4 GPU-minutes): what round 1 left unmeasured, in the order it pays.
#   1. parity of the whole tree (the round-1 tail added GPU cases that were only checked on the CPU: procedural maps)
#   2. the experimental packed-fp32 epilogue (B2C_TC_PACKED=1): parity + the tc_linear probe, A/B against the default
#   3. launch-shape sweep of the two scene-step kernels as they are now (48-register state kernel, mask-fed lidar)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
bash tools/gpu_tc_probe.sh 2>&1 | tee gpurun_out/tc_probe.log
{
echo "== default"; timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
for cfg in "2 128" "2 96" "3 96" "1 64"; do
  set -- $cfg
  echo "== state kernel: scenes per CTA $1, threads $2"
  B2C_ENV_GROUP=$1 B2C_ENV_THREADS=$2 timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
done
for cfg in "1 96" "1 160" "2 128" "2 256"; do
  set -- $cfg
  echo "== lidar kernel: scenes per CTA $1, threads $2"
  B2C_LIDAR_GROUP=$1 B2C_LIDAR_THREADS=$2 timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
done
} | tee gpurun_out/sweep.log
