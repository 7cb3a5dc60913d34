#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py -q -m gpu -x 2>&1 | tail -6
for m in 0 1; do
  echo "== B2C_ENV_SPLIT=$m"
  B2C_ENV_SPLIT=$m timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
done
for cfg in "2 128 2" "3 128 1" "3 128 3" "2 96 2" "3 128 2 256"; do
  set -- $cfg
  echo "== split: state group $1 threads $2, lidar group $3 threads ${4:-128}"
  B2C_ENV_SPLIT=1 B2C_ENV_GROUP=$1 B2C_ENV_THREADS=$2 B2C_LIDAR_GROUP=$3 B2C_LIDAR_THREADS=${4:-128} timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
done
B2C_ENV_SPLIT=1 timeout 120 python tools/env_perf.py 4096 10 parking_lot 2>&1 | tail -1
B2C_ENV_SPLIT=1 timeout 120 python tools/env_perf.py 4096 40 tollgate 2>&1 | tail -1
