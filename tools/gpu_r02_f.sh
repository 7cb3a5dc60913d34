#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_tc.log
timeout 120 python tools/fused_probe.py 2>&1 | tail -20 | tee gpurun_out/fused_probe.json
B2C_TC_PRODUCTS=3 timeout 120 python tools/fused_probe.py 2>&1 | tail -20 | tee gpurun_out/fused_probe_p3.json
timeout 120 python tools/fused_trace.py 92 2>&1 | tail -12 | tee gpurun_out/fused_trace.txt
