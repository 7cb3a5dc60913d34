#!/bin/bash
# Learner: forward of every network in the one-kernel form (b2c_tc_mlp2_train), A/B against the layer kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_learner_gpu.py tests/test_ref_golden_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_tc.log
{
echo "== fused training forward"; timeout 120 python tools/learn_time.py 65536
echo "== layer kernels"; B2C_TC_FUSED_TRAIN=0 timeout 120 python tools/learn_time.py 65536
} 2>&1 | tee gpurun_out/learn_time.log
timeout 300 python tools/train_time.py 16 6 2>&1 | tail -5 | tee gpurun_out/train_time.log
