#!/bin/bash
# Round 2, visit k: whole GPU suite, then the ncu evidence (launch list of a bench run, --set full captures of the scene
# step, the fused network, the learner kernels and the critic-obs fusion).
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --train-iters 0 --no-cpu-baseline > gpurun_out/b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"env_|tc_mlp2" -s 24 -c 3 -f -o gpurun_out/r02_step python bench.py --steps 4 --warmup 3 --train-iters 0 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_wgrad|head_backward|tc_linear|wgrad_reduce" -s 30 -c 10 -f -o gpurun_out/r02_learner python tools/learn_perf.py > gpurun_out/ncu_learner.log 2>&1; tail -2 gpurun_out/ncu_learner.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cc_obs_fuse" -s 6 -c 1 -f -o gpurun_out/r02_fuse python bench.py --config c3 --steps 4 --warmup 3 --train-iters 0 --no-cpu-baseline > gpurun_out/ncu_fuse.log 2>&1; tail -2 gpurun_out/ncu_fuse.log
ls -la gpurun_out/*.ncu-rep
