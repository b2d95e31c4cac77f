#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02m
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "frames_in_flight" > ${T}_fif_test.log 2>&1
echo "== fif test: exit $?"; tail -3 ${T}_fif_test.log; grep -E "^E " ${T}_fif_test.log | head
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_1gpu_$i.json 2> ${T}_bench_1gpu_$i.err
echo "== bench: exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/r02m_bench_1gpu_$i.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','single_frame','clocks')})
PY
tail -3 ${T}_bench_1gpu_$i.err
done
