#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step -s 20 -c 1 -f -o gpurun_out/env_step python tools/env_perf.py 4096 40 intersection > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
