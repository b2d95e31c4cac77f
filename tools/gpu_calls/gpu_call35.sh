#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02G
timeout -s KILL 420 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 150 -k "layer2_pair_window" > ${T}_tests.log 2>&1
echo "== layer2 tests (both epilogue forms): exit $?"; tail -3 ${T}_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests.log | head -20
MPX_CONV_MODE=27312139 timeout -s KILL 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_pipeline.py tests/test_zz_gpu_fullsize.py -m gpu -q --timeout 300 -k "resnet34_engine or network_with or pipeline_matches or fused_pipeline_equals or full_size" > ${T}_tests2.log 2>&1
echo "== network / pipeline / full-size tests under the staged mode: exit $?"; tail -3 ${T}_tests2.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests2.log | head -20
