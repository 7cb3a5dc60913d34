#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list and full captures of the two heavy kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --train-iters 0 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_ -s 20 -c 2 -f -o gpurun_out/env_step python bench.py --steps 10 --warmup 5 --no-cpu-baseline --train-iters 0 > gpurun_out/ncu_env.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_linear -s 21 -c 2 -f -o gpurun_out/tc_linear python bench.py --steps 10 --warmup 5 --no-cpu-baseline --train-iters 0 > gpurun_out/ncu_tc.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/learn_launches.csv python tools/learn_perf.py > /dev/null 2>&1
timeout 200 python tools/mlp_perf.py > gpurun_out/mlp_perf.json 2>&1
timeout 200 python tools/bookkeeping_perf.py > gpurun_out/bookkeeping_perf.json 2>&1
ls -la gpurun_out
