#!/bin/bash
# tc_linear iteration: tensor-core parity tests first (short timeout: a pipeline bug would hang), then the whole suite,
# bench line and the A/B against streamed weights.
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_tc_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_tc.log
grep -q "passed" gpurun_out/pytest_tc.log || { echo "tc tests did not pass: stopping"; exit 1; }
grep -q "failed\|error" gpurun_out/pytest_tc.log && { echo "tc tests failed: stopping"; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.1fM ms/step %.4f e2e %.1fM (%.3f ms) kernels %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], d["kernel_ms"]))
PY
}
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --train-iters 0 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
show gpurun_out/bench_iter.json
B2C_TC_STREAM_WEIGHTS=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --train-iters 0 > gpurun_out/bench_streamed.json 2>> gpurun_out/bench_iter.err
echo "streamed weights:"; show gpurun_out/bench_streamed.json
timeout 200 python tools/mlp_perf.py > gpurun_out/mlp_perf.json 2>&1; tail -c 1500 gpurun_out/mlp_perf.json
