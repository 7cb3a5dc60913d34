#!/bin/bash
# Scene step: dynamic scene-group scheduling in the lidar kernel (A/B against static striding); graphed meta update
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_env.log
{
echo "== dynamic"
for m in "4096 40 intersection" "4096 40 roundabout" "1024 40 tollgate" "4096 10 parking_lot"; do timeout 120 python tools/env_perf.py $m 2>&1 | tail -1; done
echo "== static"
for m in "4096 40 intersection" "1024 40 tollgate"; do B2C_LIDAR_STATIC=1 timeout 120 python tools/env_perf.py $m 2>&1 | tail -1; done
} | tee gpurun_out/env_perf.log
timeout 900 python -m pytest tests/test_learner_gpu.py tests/test_ref_golden_gpu.py tests/test_trainer_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_learner.log
timeout 300 python tools/train_time.py 16 8 2>&1 | tail -8 | tee gpurun_out/train_time.log
echo "== meta graph off"
B2C_LEARN_GRAPH=0 timeout 300 python tools/train_time.py 16 4 2>&1 | tail -2 | tee gpurun_out/train_time_nograph.log
