#!/bin/bash
# Round 2, first visit: the whole GPU suite (new: full-size oracle parity at C2-C5, global-row normalisation, plain value
# loss), then one bench line per BASELINE configuration.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
for c in c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 100 --warmup 10 --train-iters 0 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -3 gpurun_out/bench_$c.err; cat gpurun_out/bench_$c.json
done
timeout 600 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_ref_c2.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_ref_c2.json
ls -la gpurun_out
