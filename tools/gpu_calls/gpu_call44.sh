#!/bin/bash
# 4 GPUs: weak-scaling bench line (one detection per rank, two frames in flight per rank)
mkdir -p gpurun_out
T=gpurun_out/r02L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout -s KILL 400 $TR bench.py --gpus 8 --steps 20 --warmup 3 > ${T}_bench_weak8.json 2> ${T}_bench_weak8.err
echo "== weak N=8: exit $?"; grep '^{' ${T}_bench_weak8.json | cut -c1-300; tail -2 ${T}_bench_weak8.err
