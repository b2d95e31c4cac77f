#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02w
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_detector.py -m gpu -q -x -k "fused_maxpool or run_detector or pair_window64" > ${T}_tests.log 2>&1
echo "== tests: exit $?"; tail -3 ${T}_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests.log | head -20
timeout 300 python tools/gpu_pool_ab.py 576 > ${T}_pool_ab.json 2> ${T}_pool_ab.err
echo "== pool ab: exit $?"; cat ${T}_pool_ab.json; tail -3 ${T}_pool_ab.err
timeout 300 python tools/gpu_pool_ab.py 1 >> ${T}_pool_ab.json 2>> ${T}_pool_ab.err
echo "== pool ab n=1: exit $?"; tail -1 ${T}_pool_ab.json
MPX_CONV_MODE=2146315 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_fused.json 2> ${T}_bench_fused.err
echo "== bench fused: exit $?"; cut -c1-250 ${T}_bench_fused.json; grep -o '"single_frame": {[^}]*}' ${T}_bench_fused.json; grep -o '"conv_ms_per_step": [0-9.]*' ${T}_bench_fused.json; tail -3 ${T}_bench_fused.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_default.json 2> ${T}_bench_default.err
echo "== bench default: exit $?"; cut -c1-250 ${T}_bench_default.json; grep -o '"single_frame": {[^}]*}' ${T}_bench_default.json
for M in 4243467 6340619; do
MPX_CONV_MODE=$M timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_mode$M.json 2> ${T}_bench_mode$M.err
echo "== bench mode $M (bit 22 = late PDL trigger): exit $?"; cut -c1-250 ${T}_bench_mode$M.json; grep -o '"single_frame": {[^}]*}' ${T}_bench_mode$M.json; tail -2 ${T}_bench_mode$M.err
done
