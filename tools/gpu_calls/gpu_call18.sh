#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02q
timeout 1200 python tools/gpu_frames_in_flight.py --reserve 0 --slots 2,3 --priority high,low,0 --steps 24 --out ${T}_frames_in_flight.json > ${T}_fif.log 2>&1
echo "== frames in flight: exit $?"; grep reserve_sms ${T}_fif.log | cut -c1-450; tail -5 ${T}_fif.log | grep -v reserve_sms | cut -c1-300
