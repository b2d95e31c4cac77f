#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02f
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_zz_gpu_fullsize.py -m gpu -q -rs -s -k "wide or full_size" > ${T}_tests.log 2>&1
echo "== wide + fullsize tests: exit $?"; grep -E "passed|failed|^E  " ${T}_tests.log | head; grep -E "about the offset|engine-fp32" ${T}_tests.log | cut -c1-200
for k in conv_windowq conv_window2q conv_igemm2; do
  timeout 500 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:${k} --launch-count 3 -o ${T}_ncu_${k} python tools/profile_step.py --stage coarse > ${T}_ncu_${k}.log 2>&1
  echo "== ncu ${k}: exit $?"; ls -la ${T}_ncu_${k}.ncu-rep 2>/dev/null | cut -c20-
done
