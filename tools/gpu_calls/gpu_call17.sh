#!/bin/bash
mkdir -p gpurun_out
for c in 8 32; do
echo "== CUDA_DEVICE_MAX_CONNECTIONS=$c"
CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 600 python tools/gpu_e2e_variants.py 2>&1 | grep ms_per_frame | sed -n '1p;4p;6p' | cut -c1-200
done
