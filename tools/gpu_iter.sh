#!/bin/bash
# Kernel iteration visit: parity, a small launch-shape sweep, one full ncu capture of the default shape.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for cfg in "4 256" "2 128" "1 64" "1 128" "1 96" "2 256" "2 192" "3 192"; do
  set -- $cfg
  echo "== group $1 threads $2"
  B2C_ENV_GROUP=$1 B2C_ENV_THREADS=$2 timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
done | tee gpurun_out/sweep.log
timeout 120 python tools/env_perf.py 4096 10 parking_lot 2>&1 | tail -1 | tee -a gpurun_out/sweep.log
timeout 120 python tools/env_perf.py 4096 40 tollgate 2>&1 | tail -1 | tee -a gpurun_out/sweep.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step -s 20 -c 1 -f -o gpurun_out/env_step python tools/env_perf.py 4096 40 intersection > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
