#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_env.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --train-iters 0 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_iter.json"))
print("value %.1fM ms/step %.4f e2e %s kernels %s" % (d["value"] / 1e6, d["ms_per_step"], {k: v for k, v in d["e2e"].items() if k != "api"}, d["kernel_ms"]))
PY
timeout 200 python tools/tc_probe.py 2>&1 | tee gpurun_out/tc_probe.json
