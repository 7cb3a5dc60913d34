#!/bin/bash
# Scene step: own-lane lasers + table-driven spread (LIDAR_OWN variants built as variants/lib_own*.so)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_env.log
{
echo "== own2 (default)"
for m in "4096 40 intersection" "4096 40 roundabout" "1024 40 tollgate" "4096 10 parking_lot"; do timeout 120 python tools/env_perf.py $m 2>&1 | tail -1; done
cp copo_b200/libcopo_b200.so /tmp/lib_default.so
for v in 0 1 3; do
  echo "== own$v"; cp variants/lib_own$v.so copo_b200/libcopo_b200.so
  for m in "4096 40 intersection" "1024 40 tollgate" "4096 10 parking_lot"; do timeout 120 python tools/env_perf.py $m 2>&1 | tail -1; done
done
cp /tmp/lib_default.so copo_b200/libcopo_b200.so
} | tee gpurun_out/env_perf.log
