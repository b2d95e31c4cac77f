#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02k
timeout 1200 python tools/gpu_frames_in_flight.py --reserve 0,16 --slots 2,3 --priority 0,1 --steps 20 --out ${T}_frames_in_flight.json > ${T}_fif.log 2>&1
echo "== frames in flight: exit $?"; grep reserve_sms ${T}_fif.log | cut -c1-420; tail -5 ${T}_fif.log | grep -v reserve_sms | cut -c1-300
