#!/bin/bash
# Training iterations at the C4 per-GPU shape (1024 scenes x 40 agents, T = 16) on N GPUs, minibatch weak-scaled.
mkdir -p gpurun_out
N=${1:-8}
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --config c4 --gpus $N --steps 20 --warmup 5 --train-iters 4 > gpurun_out/bench_c4_train_n$N.json 2> gpurun_out/bench_c4_train_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_c4_train_n$N.json').read().strip().splitlines()[-1]); print('N=$N c4 value %.1fM e2e %.1fM'%(d['value']/1e6, d['e2e']['value']/1e6)); print(json.dumps(d['train'])[:1000])"
timeout 300 python bench.py --config c4 --steps 20 --warmup 5 --train-iters 4 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_c4_train_n1.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_c4_train_n1.json').read().strip().splitlines()[-1]); print('N=1 c4 value %.1fM'%(d['value']/1e6)); print(json.dumps(d['train'])[:1000])"
