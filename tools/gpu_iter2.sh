#!/bin/bash
# Kernel iteration: whole GPU parity suite, bench line (no CPU / training legs), A/B of the epilogue store width,
# scene-step timings per map, source-level capture of the env kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.1fM ms/step %.4f e2e %.1fM (%.3f ms) kernels %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], d["kernel_ms"]))
PY
}
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --train-iters 0 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
show gpurun_out/bench_iter.json
B2C_TC_NARROW_STORES=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --train-iters 0 > gpurun_out/bench_narrow.json 2>> gpurun_out/bench_iter.err
echo "narrow stores:"; show gpurun_out/bench_narrow.json
timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
timeout 120 python tools/env_perf.py 4096 40 tollgate 2>&1 | tail -1
timeout 120 python tools/env_perf.py 4096 10 parking_lot 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_ -s 20 -c 2 -f -o gpurun_out/env_step_g2 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --train-iters 0 > gpurun_out/ncu_env.log 2>&1
tail -1 gpurun_out/ncu_env.log
