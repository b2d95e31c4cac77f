#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02j
timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q -k "msaa or antialias" > ${T}_msaa.log 2>&1
echo "== msaa tests: exit $?"; tail -3 ${T}_msaa.log
timeout 900 python tools/gpu_frames_in_flight.py --reserve 0,8,16,24,32 --steps 20 --out ${T}_frames_in_flight.json > ${T}_fif.log 2>&1
echo "== frames in flight: exit $?"; grep reserve_sms ${T}_fif.log; tail -5 ${T}_fif.log | cut -c1-300
