#!/bin/bash
mkdir -p gpurun_out
for epi in 8 12 16; do
  echo "=== epilogue warps $epi"
  B2C_TC_EPI_WARPS=$epi timeout 120 python -m pytest tests/test_tc_gpu.py -m gpu -x -q 2>&1 | tail -2
  B2C_TC_EPI_WARPS=$epi timeout 100 python tools/tc_probe.py > gpurun_out/tc_probe_epi$epi.json 2>&1
  python - gpurun_out/tc_probe_epi$epi.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
for k, v in d.items():
    print(k, {n: round(x, 4) for n, x in v.items()})
PY
done
echo "=== 3 products (epilogue warps 8)"
B2C_TC_PRODUCTS=3 timeout 120 python -m pytest tests/test_tc_gpu.py -m gpu -q 2>&1 | tail -5
B2C_TC_PRODUCTS=3 timeout 100 python tools/tc_probe.py > gpurun_out/tc_probe_p3.json 2>&1
python - gpurun_out/tc_probe_p3.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
for k, v in d.items():
    print(k, {n: round(x, 4) for n, x in v.items()})
PY
echo "=== packed fp32x2 epilogue math (experimental, B2C_TC_PACKED=1)"
B2C_TC_PACKED=1 timeout 120 python -m pytest tests/test_tc_gpu.py -m gpu -q 2>&1 | tail -5
B2C_TC_PACKED=1 timeout 100 python tools/tc_probe.py > gpurun_out/tc_probe_packed.json 2>&1
python - gpurun_out/tc_probe_packed.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
for k, v in d.items():
    print(k, {n: round(x, 4) for n, x in v.items()})
PY
