#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02B
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 > ${T}_gpu_tests.log 2>&1
echo "== gpu tests: exit $?"; tail -3 ${T}_gpu_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_gpu_tests.log | head -20
