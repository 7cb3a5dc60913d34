#!/bin/bash
# Round 2, visit g: whole GPU suite after the fused-network / LDS changes, bench lines c2..c5.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
for c in c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 100 --warmup 10 --train-iters 0 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -3 gpurun_out/bench_$c.err
done
python - <<'PY'
import json
for c in ("c2","c3","c4","c5"):
    try:
        d=json.load(open(f"gpurun_out/bench_{c}.json"))
        print(c, "value %.1fM"%(d["value"]/1e6), "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.1fM"%(d["e2e"]["value"]/1e6), d["kernel_ms"], d.get("train") and d["train"]["agent_env_steps_per_s_per_gpu"])
    except Exception as e: print(c, "failed", e)
PY
