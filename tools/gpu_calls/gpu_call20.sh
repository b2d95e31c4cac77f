#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02s
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x > ${T}_net_tests.log 2>&1
echo "== net tests: exit $?"; tail -3 ${T}_net_tests.log; grep -E "^E " ${T}_net_tests.log | head
timeout 600 python tools/gpu_layer_table.py --mpx-only --conv-modes 1097739 --out ${T}_layer_table.json > ${T}_layer_table.log 2>&1
echo "== layer table: exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s_layer_table.json'))
for r in d['rows']: print(r['layer'], r['count'], round(r['mpx_ms'],3), round(r['mpx_mode1097739_ms'],3), round(r['mpx_tflops']))
print(d['total'])
PY
timeout 900 python -m pytest tests/test_zz_gpu_fullsize.py tests/test_gpu_pipeline.py -m gpu -q -x > ${T}_pipe_tests.log 2>&1
echo "== pipeline + fullsize tests: exit $?"; tail -3 ${T}_pipe_tests.log; grep -E "^E " ${T}_pipe_tests.log | head
