#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_env.log
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.1fM ms/step %.4f e2e %.1fM (%.3f ms) kernels %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], d["kernel_ms"]))
PY
}
for ch in 4 8 2; do
  B2C_HOST_CHUNKS=$ch timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --train-iters 0 > gpurun_out/bench_c$ch.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
  echo "host chunks $ch:"; show gpurun_out/bench_c$ch.json
done
timeout 120 python tools/env_perf.py 4096 40 intersection 2>&1 | tail -1
timeout 120 python tools/env_perf.py 4096 40 tollgate 2>&1 | tail -1
timeout 200 python tools/tc_probe.py 2>&1 | tee gpurun_out/tc_probe.json
