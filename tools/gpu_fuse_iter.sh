#!/bin/bash
# Critic-obs fusion with the operand output: parity (learner + golden + trajectory + trainer tests), bookkeeping
# bandwidths, C3 bench line, training-iteration phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_learner_gpu.py tests/test_ref_golden_gpu.py tests/test_trajectory_gpu.py tests/test_trainer_gpu.py tests/test_compat_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_learner.log
timeout 200 python tools/bookkeeping_perf.py 2>&1 | tee gpurun_out/bookkeeping_perf.json | grep -A4 cc_obs_fuse
timeout 600 python bench.py --config c3 --steps 100 --warmup 10 --train-iters 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 value %.1fM ms %.4f'%(d['value']/1e6, d['ms_per_step']), d['kernel_ms'])"
timeout 300 python tools/train_time.py 16 5 2>&1 | tail -3
