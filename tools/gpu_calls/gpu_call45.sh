#!/bin/bash
# 2 GPUs: head-done event recorded before the coarse all-gather -- multi-rank parity test and the weak line
mkdir -p gpurun_out
T=gpurun_out/r02M
timeout -s KILL 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 200 > ${T}_multi_tests.log 2>&1
echo "== multi tests: exit $?"; tail -2 ${T}_multi_tests.log | cut -c1-200
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523"
timeout -s KILL 300 $TR bench.py --gpus 2 --steps 30 --warmup 3 > ${T}_bench_weak2.json 2> ${T}_bench_weak2.err
echo "== weak N=2: exit $?"; grep '^{' ${T}_bench_weak2.json | cut -c1-260
