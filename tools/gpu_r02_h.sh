#!/bin/bash
# Round 2, visit h: learner without fp32 intermediates (bias gradients from the ones column / the head kernel).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_learner_gpu.py tests/test_tc_gpu.py tests/test_trainer_gpu.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_learner.log
timeout 120 python tools/learn_time.py 65536 2>&1 | tail -1 | tee gpurun_out/learn_time.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/learn_launches.csv python tools/learn_perf.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/learn_launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4][:70]; agg[name][0] += 1; agg[name][1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%-70s %3d  %7.1f us  %4.1f%%" % (k, v[0], v[1] / v[0] / (1000 if v[1] / v[0] > 5000 else 1), 100 * v[1] / tot))
print("total per call (3 calls):", tot / 3)
PY
