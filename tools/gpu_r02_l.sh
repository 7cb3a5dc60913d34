#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/pytest_gpu.log
timeout 120 python tools/learn_time.py 65536 2>&1 | tail -1 | tee gpurun_out/learn_time.json
timeout 120 python tools/bookkeeping_perf.py 2>&1 | tail -12 | tee gpurun_out/bookkeeping_perf.json
