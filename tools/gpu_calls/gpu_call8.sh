#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02h
timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_multi.py -m gpu -q -rs -s > ${T}_tests.log 2>&1
echo "== icp + multi tests: exit $?"; grep -E "passed|failed|^E  |object [0-9]" ${T}_tests.log | head -20
for sc in weak strong; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --scaling ${sc} > ${T}_bench_2gpu_${sc}.json 2> ${T}_bench_2gpu_${sc}.err
  echo "== bench 2 GPUs ${sc}: exit $?"; cut -c1-260 ${T}_bench_2gpu_${sc}.json; tail -2 ${T}_bench_2gpu_${sc}.err | cut -c1-300
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --workload ycbv21 > ${T}_bench_2gpu_ycbv21.json 2> ${T}_bench_2gpu_ycbv21.err
echo "== bench 2 GPUs ycbv21 strong: exit $?"; cut -c1-260 ${T}_bench_2gpu_ycbv21.json; tail -2 ${T}_bench_2gpu_ycbv21.err | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --scaling strong --workload ycbv21 > ${T}_bench_1gpu_ycbv21.json 2> ${T}_bench_1gpu_ycbv21.err
echo "== bench 1 GPU ycbv21: exit $?"; cut -c1-260 ${T}_bench_1gpu_ycbv21.json; tail -2 ${T}_bench_1gpu_ycbv21.err | cut -c1-300
