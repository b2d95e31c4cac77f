#!/bin/bash
# First GPU call of a round: everything that was written without a GPU at hand, then the A/B of the experimental bits.
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/gpu_first_call.sh'
# Every step is its own process with its own timeout: a trap in one experimental kernel (bounded mbarrier waits trap after
# ~4 s) poisons that process's CUDA context only.  Logs land in gpurun_out/.
mkdir -p gpurun_out
export MPX_EXPERIMENTAL=1
run() {  # name, pytest args...
  local name=$1; shift
  timeout 420 python -m pytest "$@" -q -x > "gpurun_out/first_${name}.log" 2>&1
  echo "== ${name}: exit $?  $(tail -1 gpurun_out/first_${name}.log)"
}
run fullsize        tests/test_zz_gpu_fullsize.py -m gpu
run runner          tests/test_prediction_runner.py -m gpu
run observers       tests/test_gpu_net.py -k "experimental_window_observers"
run pairwin_l34     tests/test_gpu_net.py -k "experimental_pair_window_kernel and (l3 or l4 or odd_size) and not l2"
run pairwin_l2_128  tests/test_gpu_net.py -k "experimental_pair_window_kernel and l2_128wide"
run pairs_l2        tests/test_gpu_net.py -k "experimental_pair_window_kernel and l2_pairs"
run pairs_64        tests/test_gpu_net.py -k "experimental_pair_window64"
unset MPX_EXPERIMENTAL
# 11 default | 2048 gated refills | 4096 layer3/4 pair-window | 8192 layer2 128-wide pair-window | 16384 layer2 pairs |
# 32768 stem+layer1 pairs | 49152 both pair kernels | +65536 residual preload
timeout 1000 python tools/gpu_ab.py --conv 11,2059,4107,8203,16395,32779,49163,114699 --steps 20 --rounds 2 --timeout 150 \
  --out gpurun_out/ab_modes.json > gpurun_out/ab_modes.log 2>&1
echo "== A/B: exit $?"; tail -90 gpurun_out/ab_modes.log
