#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02r
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_windowq --launch-skip 1 --launch-count 2 -o ${T}_windowq -f python tools/profile_step.py --stage coarse > ${T}_ncu.log 2>&1
echo "== ncu exit $?"; tail -3 ${T}_ncu.log; ls -la ${T}_windowq.ncu-rep
