#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_env_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_ -s 40 -c 2 -f -o gpurun_out/env_split python tools/env_perf.py 4096 40 intersection > gpurun_out/ncu_split.log 2>&1
tail -2 gpurun_out/ncu_split.log
