#!/bin/bash
# new schedule options: parity tests, A/B timings, bench lines per option
mkdir -p gpurun_out
T=gpurun_out/r02w
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_detector.py tests/test_gpu_kernels.py -m gpu -q -k "fused_maxpool or chunked or run_detector or pair_window64 or raster or textured" > ${T}_tests.log 2>&1
echo "== tests: exit $?"; tail -3 ${T}_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests.log | head -20
timeout 600 python tools/gpu_pool_ab.py 576 0,24,32,48,64,96 > ${T}_pool_ab.json 2> ${T}_pool_ab.err
echo "== pool/chunk ab: exit $?"; cat ${T}_pool_ab.json; tail -3 ${T}_pool_ab.err
timeout 300 python tools/gpu_pool_ab.py 1 >> ${T}_pool_ab.json 2>> ${T}_pool_ab.err
echo "== pool ab n=1: exit $?"; tail -1 ${T}_pool_ab.json
run_bench() {  # name, conv mode, chunk
  MPX_CONV_MODE=$2 MPX_NET_CHUNK=$3 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_$1.json 2> ${T}_bench_$1.err
  echo "== bench $1 (mode $2, chunk $3): exit $?"; cut -c1-200 ${T}_bench_$1.json; grep -o '"single_frame": {[^}]*}' ${T}_bench_$1.json; grep -o '"conv_ms_per_step": [0-9.]*' ${T}_bench_$1.json; tail -2 ${T}_bench_$1.err
}
run_bench default 49163 0
run_bench fusedpool 2146315 0
run_bench latepdl 4243467 0
run_bench pool_pdl 6340619 0
run_bench pool_pdl_chunk48 6340619 48
run_bench pool_pdl_chunk32 6340619 32
run_bench chunk48 49163 48
