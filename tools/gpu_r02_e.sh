#!/bin/bash
# Round 2, visit e: packed epilogue math everywhere (new clamp / reciprocal in tanh), fused kernel v2 timing + ncu.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_learner_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_tc.log
timeout 120 python tools/fused_probe.py 2>&1 | tail -20 | tee gpurun_out/fused_probe.json
B2C_TC_PRODUCTS=3 timeout 120 python tools/fused_probe.py 2>&1 | tail -20 | tee gpurun_out/fused_probe_p3.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_mlp2 -s 3 -c 1 -f -o gpurun_out/tc_mlp2 python tools/fused_probe.py > gpurun_out/ncu_mlp2.log 2>&1
tail -3 gpurun_out/ncu_mlp2.log
