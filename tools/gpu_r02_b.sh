#!/bin/bash
# Round 2, visit b: programmatic dependent launch of the scene-step kernels (on / off), the fixed per-trajectory test,
# and the packed-fp32 epilogue A/B that round 1 left unmeasured.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_trajectory_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_env.log
{
for m in "4096 40 intersection" "4096 40 roundabout" "1024 40 tollgate"; do
  echo "== pdl $m"; timeout 120 python tools/env_perf.py $m 2>&1 | tail -1
  echo "== no pdl $m"; B2C_ENV_PDL=0 timeout 120 python tools/env_perf.py $m 2>&1 | tail -1
done
} | tee gpurun_out/env_perf_pdl.log
{
echo "=== default epilogue"
timeout 100 python tools/tc_probe.py 2>&1 | tee gpurun_out/tc_probe_default.json | python -c "import json,sys; d=json.load(sys.stdin); [print(k, {n: round(x, 4) for n, x in v.items()}) for k, v in d.items()]"
for epi in 8 16; do
echo "=== packed fp32x2 epilogue math, $epi epilogue warps"
B2C_TC_EPI_WARPS=$epi B2C_TC_PACKED=1 timeout 200 python -m pytest tests/test_tc_gpu.py tests/test_learner_gpu.py -m gpu -q 2>&1 | tail -5
B2C_TC_EPI_WARPS=$epi B2C_TC_PACKED=1 timeout 100 python tools/tc_probe.py 2>&1 | tee gpurun_out/tc_probe_packed_epi$epi.json | python -c "import json,sys; d=json.load(sys.stdin); [print(k, {n: round(x, 4) for n, x in v.items()}) for k, v in d.items()]"
done
} 2>&1 | tee gpurun_out/tc_packed.log
