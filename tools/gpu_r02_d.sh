#!/bin/bash
# Round 2, visit d: one-kernel network parity (tolerances), trajectory test, ncu source-level capture of the fused kernel.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_fused.log
timeout 600 python -m pytest tests/test_trajectory_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_traj.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_mlp2 -s 3 -c 1 -f -o gpurun_out/tc_mlp2 python tools/fused_probe.py > gpurun_out/ncu_mlp2.log 2>&1
tail -3 gpurun_out/ncu_mlp2.log
