#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_env.log
for m in "4096 40 intersection" "4096 40 roundabout" "1024 40 tollgate" "4096 10 parking_lot"; do timeout 120 python tools/env_perf.py $m 2>&1 | tail -1; done | tee gpurun_out/env_perf.log
timeout 300 python bench.py --steps 50 --warmup 10 --train-iters 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c2 value %.1fM ms %.4f frac %.3f'%(d['value']/1e6, d['ms_per_step'], d['roofline']['frac']), d['kernel_ms'])"
