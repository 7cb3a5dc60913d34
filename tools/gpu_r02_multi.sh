#!/bin/bash
# Round 2, multi-GPU visit: bench (both arms) under torchrun, training iterations at the C4 per-GPU shape with the
# all-reduce share, identical-parameters check, host-link ceiling.
mkdir -p gpurun_out
N=${1:-2}
run() { timeout ${T:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
NCCL_DEBUG=INFO run 29611 bench.py --gpus $N --steps 50 --warmup 10 --train-iters 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -c "NCCL INFO" gpurun_out/bench_n$N.err; grep -m3 "NVLS\|nranks\|Connected all" gpurun_out/bench_n$N.err | cut -c1-200; tail -2 gpurun_out/bench_n$N.err | cut -c1-300
cut -c1-600 gpurun_out/bench_n$N.json
run 29614 bench.py --config c4 --gpus $N --steps 50 --warmup 10 --train-iters 3 --train-fragment 32 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; tail -2 gpurun_out/bench_c4_n$N.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_n$N.json')); print('c4 value %.1fM e2e %.1fM'%(d['value']/1e6, d['e2e']['value']/1e6)); print(json.dumps(d['train'])[:900])"
run 29612 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err; cut -c1-200 gpurun_out/bench_ref_n$N.json
run 29613 tools/train_check.py 2>&1 | tail -6 | tee gpurun_out/train_check_n$N.log
run 29615 tools/pcie_ceiling.py 2>&1 | tail -1 | tee gpurun_out/pcie_ceiling_n$N.json
