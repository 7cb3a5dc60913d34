#!/bin/bash
# One GPU visit: parity suite, smoke, scene-step timings over a few launch shapes, bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
{
echo "== default"; timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
for cfg in "7 128" "10 128" "64 256" "64 64" "64 96"; do
  set -- $cfg
  echo "== lidar ctas/SM $1 threads $2"
  B2C_LIDAR_CTAS=$1 B2C_LIDAR_THREADS=$2 timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
done
echo "== state group 2"; B2C_ENV_GROUP=2 timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
echo "== fused"; B2C_ENV_SPLIT=0 timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
timeout 120 python tools/env_perf.py 4096 10 parking_lot 2>&1 | tail -1
timeout 120 python tools/env_perf.py 4096 40 tollgate 2>&1 | tail -1
} | tee gpurun_out/sweep.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
