#!/bin/bash
# Round 2, visit c: first run of the one-kernel network (parity under a short timeout, then timing), the trajectory test,
# PDL variants of the scene step.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -x -q -k one_kernel 2>&1 | tail -15 | tee gpurun_out/pytest_fused.log
timeout 120 python tools/fused_probe.py 2>&1 | tail -20 | tee gpurun_out/fused_probe.json
B2C_TC_PRODUCTS=3 timeout 120 python tools/fused_probe.py 2>&1 | tail -20 | tee gpurun_out/fused_probe_p3.json
timeout 600 python -m pytest tests/test_trajectory_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_traj.log
{
for p in 0 1 2 3; do
  echo "== B2C_ENV_PDL=$p"; B2C_ENV_PDL=$p timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
done
} | tee gpurun_out/env_perf_pdl2.log
