#!/bin/bash
# Scene-step iteration (round 2): env parity, scene step timing (specialised vs generic kernels), source-level ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_env.log
{
for m in "4096 40 intersection" "4096 40 roundabout" "1024 40 tollgate" "4096 10 parking_lot"; do
  echo "== specialised $m"; timeout 120 python tools/env_perf.py $m 2>&1 | tail -1
  echo "== generic $m"; B2C_ENV_GENERIC=1 timeout 120 python tools/env_perf.py $m 2>&1 | tail -1
done
} | tee gpurun_out/env_perf.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_ -s 60 -c 2 -f -o gpurun_out/env_step python tools/env_perf.py 4096 40 intersection > gpurun_out/ncu_env.log 2>&1
tail -2 gpurun_out/ncu_env.log
