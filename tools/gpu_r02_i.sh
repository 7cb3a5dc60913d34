#!/bin/bash
# Round 2, visit i: CUDA-graph capture of the minibatch gradient step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_learner_gpu.py tests/test_trainer_gpu.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_learner.log
echo "graph:"; timeout 120 python tools/learn_time.py 65536 2>&1 | tail -2 | tee gpurun_out/learn_time_graph.json
echo "eager:"; B2C_LEARN_GRAPH=0 timeout 120 python tools/learn_time.py 65536 2>&1 | tail -1 | tee gpurun_out/learn_time_eager.json
timeout 600 python bench.py --steps 20 --warmup 5 --train-iters 3 --no-cpu-baseline 2> gpurun_out/bench_train.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['train'])[:700])" | tee gpurun_out/bench_train.json
tail -3 gpurun_out/bench_train.err
