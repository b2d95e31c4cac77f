#!/bin/bash
# 4 GPUs: weak-scaling bench line (one detection per rank, two frames in flight per rank)
mkdir -p gpurun_out
T=gpurun_out/r02L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519"
timeout -s KILL 400 $TR bench.py --gpus 4 --steps 20 --warmup 3 > ${T}_bench_weak4.json 2> ${T}_bench_weak4.err
echo "== weak N=4: exit $?"; grep '^{' ${T}_bench_weak4.json | cut -c1-300; tail -2 ${T}_bench_weak4.err
