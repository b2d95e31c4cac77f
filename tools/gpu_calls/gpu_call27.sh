#!/bin/bash
# ncu --set full of the 64 -> 64 pair window kernels (stem + first layer1 launches): sliding (bit 23) and reloading windows
mkdir -p gpurun_out
T=gpurun_out/r02z
MPX_CONV_MODE=10534923 timeout -s KILL 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_windows_kernel --launch-count 3 -o ${T}_windows -f python tools/profile_step.py --stage coarse > ${T}_ncu1.log 2>&1
echo "== ncu sliding exit $?"; tail -2 ${T}_ncu1.log; ls -la ${T}_windows.ncu-rep
MPX_CONV_MODE=2146315 timeout -s KILL 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_windowq_kernel --launch-count 3 -o ${T}_windowq -f python tools/profile_step.py --stage coarse > ${T}_ncu2.log 2>&1
echo "== ncu windowq exit $?"; tail -2 ${T}_ncu2.log; ls -la ${T}_windowq.ncu-rep
timeout -s KILL 300 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 150 -k "resnet34_engine or wide_resnet" > ${T}_tests.log 2>&1
echo "== engine tests (emulated oracle without TF32): exit $?"; tail -3 ${T}_tests.log | cut -c1-300
