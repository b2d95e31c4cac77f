#!/bin/bash
# ncu --set full (source-level) of the tiled rasteriser in the coarse stage, and of the stem launch of the sliding-window kernel
mkdir -p gpurun_out
T=gpurun_out/r02J
timeout -s KILL 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:raster_tiled_kernel --launch-count 1 -o ${T}_raster -f python tools/profile_step.py --stage coarse > ${T}_ncu1.log 2>&1
echo "== ncu raster exit $?"; tail -2 ${T}_ncu1.log; ls -la ${T}_raster.ncu-rep
timeout -s KILL 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_windows_kernel --launch-count 1 -o ${T}_stem -f python tools/profile_step.py --stage coarse > ${T}_ncu2.log 2>&1
echo "== ncu stem exit $?"; tail -2 ${T}_ncu2.log; ls -la ${T}_stem.ncu-rep
