#!/bin/bash
# Scene-step kernel iteration: parity of the env kernels, bench line without the CPU / training legs, source-level capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_env.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --train-iters 0 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_iter.json"))
print("value %.1fM ms/step %.4f e2e %.1fM kernels %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["kernel_ms"]))
PY
timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
timeout 120 python tools/env_perf.py 4096 40 tollgate 2>&1 | tail -1
timeout 120 python tools/env_perf.py 4096 10 parking_lot 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_ -s 20 -c 2 -f -o gpurun_out/env_step python bench.py --steps 10 --warmup 5 --no-cpu-baseline --train-iters 0 > gpurun_out/ncu_env.log 2>&1
tail -1 gpurun_out/ncu_env.log
