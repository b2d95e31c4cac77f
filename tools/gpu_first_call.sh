#!/bin/bash
# First GPU call of a round: everything that was written without a GPU at hand, then the A/B of the experimental bits.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# Every step has its own timeout (a trap in an experimental kernel ends that step, not the call); logs land in gpurun_out/.
mkdir -p gpurun_out
export MPX_EXPERIMENTAL=1
timeout 600 python -m pytest tests/test_zz_gpu_fullsize.py tests/test_prediction_runner.py -m gpu -q > gpurun_out/first_new_tests.log 2>&1
echo "new tests: exit $?"; tail -3 gpurun_out/first_new_tests.log
timeout 600 python -m pytest tests/test_gpu_net.py -k experimental -q > gpurun_out/first_experimental.log 2>&1
echo "experimental kernels: exit $?"; tail -5 gpurun_out/first_experimental.log
unset MPX_EXPERIMENTAL
timeout 900 python tools/gpu_ab.py --conv 11,2059,4107,8203,16395,32779,49163,114699 --steps 20 --rounds 3 --timeout 200 --out gpurun_out/ab_modes.json > gpurun_out/ab_modes.log 2>&1
echo "A/B: exit $?"; tail -40 gpurun_out/ab_modes.log
