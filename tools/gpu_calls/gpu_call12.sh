#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02l
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "frames_in_flight" > ${T}_fif_test.log 2>&1
echo "== fif test: exit $?"; tail -3 ${T}_fif_test.log; grep -E "^E " ${T}_fif_test.log | head
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_1gpu.json 2> ${T}_bench_1gpu.err
echo "== bench: exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','single_frame','frames_in_flight','gpu_launches','clocks')})
print(d['roofline']['frac'], d['roofline']['conv_ms_per_step'])
PY
tail -3 ${T}_bench_1gpu.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline --render-size 224x224 > ${T}_bench_1gpu_224.json 2> ${T}_bench_1gpu_224.err
echo "== bench 224: exit $?"; cut -c1-300 ${T}_bench_1gpu_224.json
