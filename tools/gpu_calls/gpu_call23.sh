#!/bin/bash
# 2 GPUs: multi-rank tests, weak and strong bench lines with two frames in flight
mkdir -p gpurun_out
T=gpurun_out/r02v
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_prediction_runner.py -m gpu -q > ${T}_multi_tests.log 2>&1
echo "== multi tests: exit $?"; tail -3 ${T}_multi_tests.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > ${T}_bench_weak.json 2> ${T}_bench_weak.err
echo "== weak: exit $?"; cut -c1-400 ${T}_bench_weak.json; tail -3 ${T}_bench_weak.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --frames-in-flight 1 > ${T}_bench_weak_fif1.json 2> ${T}_bench_weak_fif1.err
echo "== weak fif1: exit $?"; cut -c1-400 ${T}_bench_weak_fif1.json; tail -3 ${T}_bench_weak_fif1.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > ${T}_bench_strong.json 2> ${T}_bench_strong.err
echo "== strong: exit $?"; cut -c1-400 ${T}_bench_strong.json; tail -3 ${T}_bench_strong.err
