#!/bin/bash
# re-entry check: full GPU suite on HEAD, bench line, layer table
mkdir -p gpurun_out
T=gpurun_out/r02u
timeout 1500 python -m pytest tests -m gpu -q > ${T}_gpu_tests.log 2>&1
echo "== gpu tests: exit $?"; tail -4 ${T}_gpu_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_gpu_tests.log | head -20
timeout 900 python bench.py --steps 20 --warmup 3 > ${T}_bench.json 2> ${T}_bench.err
echo "== bench: exit $?"; cut -c1-1500 ${T}_bench.json; tail -3 ${T}_bench.err
timeout 600 python tools/gpu_layer_table.py --mpx-only --out ${T}_layer_table.json > ${T}_layer_table.log 2>&1
echo "== layer table: exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02u_layer_table.json'))
for r in d['rows']: print(r['layer'], r['count'], round(r['mpx_ms'],3), round(r['mpx_tflops']))
print(d.get('total'))
PY
