#!/bin/bash
# State kernel: whole groups up to a multiple of the SM count + single scenes (A/B), torch bridge test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_bridge_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_env.log
{
echo "== singles"
for m in "4096 40 intersection" "4096 40 roundabout" "1024 40 tollgate" "4096 10 parking_lot"; do timeout 120 python tools/env_perf.py $m 2>&1 | tail -1; done
echo "== whole groups only"
for m in "4096 40 intersection" "1024 40 tollgate"; do B2C_ENV_NO_SINGLES=1 timeout 120 python tools/env_perf.py $m 2>&1 | tail -1; done
} | tee gpurun_out/env_perf.log
