#!/bin/bash
# Round 2, final visit (tag r02_e): whole GPU suite, smoke, bench lines c2..c5 + reference arm, then the ncu evidence of
# the same build (launch list of a bench run, --set full captures of the scene step + fused network and of the learner).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r02_e_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_e_bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err
for c in c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 100 --warmup 10 --train-iters 0 > gpurun_out/r02_e_bench_$c.json 2> gpurun_out/bench_$c.err; tail -2 gpurun_out/bench_$c.err
done
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_e_bench_reference_c2.json 2>> gpurun_out/bench_c2.err
python - <<'PY'
import json
for c in ("c2","c3","c4","c5"):
    try:
        d=json.load(open(f"gpurun_out/r02_e_bench_{c}.json"))
        print(c, "value %.1fM"%(d["value"]/1e6), "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.1fM"%(d["e2e"]["value"]/1e6), d["kernel_ms"], d.get("train") and d["train"]["agent_env_steps_per_s_per_gpu"])
    except Exception as e: print(c, "failed", e)
PY
timeout 120 python tools/learn_time.py 65536 2>&1 | tail -1 | tee gpurun_out/r02_e_learn_time.json
timeout 300 python tools/train_time.py 16 6 2>&1 | tail -5 | tee gpurun_out/r02_e_train_time.log
for m in "4096 40 intersection" "4096 40 roundabout" "1024 40 tollgate" "4096 10 parking_lot"; do timeout 120 python tools/env_perf.py $m 2>&1 | tail -1; done | tee gpurun_out/r02_e_env_perf.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_e_launches.csv python bench.py --steps 4 --warmup 3 --train-iters 0 --no-cpu-baseline > gpurun_out/b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"env_|tc_mlp2" -s 24 -c 3 -f -o gpurun_out/r02_e_step python bench.py --steps 4 --warmup 3 --train-iters 0 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_mlp2|tc_wgrad|head_backward|tc_linear|wgrad_reduce" -s 30 -c 10 -f -o gpurun_out/r02_e_learner python tools/learn_perf.py > gpurun_out/ncu_learner.log 2>&1; tail -2 gpurun_out/ncu_learner.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cc_obs_fuse" -s 6 -c 1 -f -o gpurun_out/r02_e_fuse python bench.py --config c3 --steps 4 --warmup 3 --train-iters 0 --no-cpu-baseline > gpurun_out/ncu_fuse.log 2>&1; tail -2 gpurun_out/ncu_fuse.log
timeout 200 python tools/bookkeeping_perf.py > gpurun_out/r02_e_bookkeeping_perf.json 2>&1
ls -la gpurun_out/r02_e_*
