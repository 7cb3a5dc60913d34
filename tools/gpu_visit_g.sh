#!/bin/bash
# Round 1 version g visit: parity suite, smoke, bench line, launch list, source-level captures of the two env kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --train-iters 0 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_ -s 20 -c 2 -f -o gpurun_out/env_step python bench.py --steps 10 --warmup 5 --no-cpu-baseline --train-iters 0 > gpurun_out/ncu_env.log 2>&1
tail -2 gpurun_out/ncu_env.log
ls -la gpurun_out
