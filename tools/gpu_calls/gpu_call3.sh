#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02e
timeout 900 python -m pytest tests -m gpu -q -rs --deselect tests/test_zz_gpu_fullsize.py > ${T}_gputests.log 2>&1
echo "== gpu tests (without full size): exit $?"; grep -E "passed|failed|error" ${T}_gputests.log | tail -3; grep -E "^E  |FAILED|Error" ${T}_gputests.log | head -40
timeout 900 python -m pytest tests/test_zz_gpu_fullsize.py -m gpu -q -rs -s > ${T}_fullsize.log 2>&1
echo "== full size: exit $?"; grep -E "passed|failed|^E  " ${T}_fullsize.log | head -20; grep -E "coarse logits|survivors|iteration 5" ${T}_fullsize.log
timeout 300 python tools/gpu_mma_probe.py ${T}_mma_probe.json > ${T}_mma_probe.log 2>&1
echo "== mma probe: exit $?"; cat ${T}_mma_probe.log | cut -c1-200
timeout 600 python tools/gpu_layer_table.py --mpx-only --conv-modes 11 --out ${T}_layer_modes.json > ${T}_layer_modes.log 2>&1
echo "== layer modes (default 49163 | 180235 = no zero-slice skip | 11 = single-CTA window kernels): exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_layer_modes.json'))
for r in d['rows']:
    print(r['layer'], r['count'], ' '.join(f"{k[4:-3]}={v:.3f}" for k,v in r.items() if k.startswith('mpx_') and k.endswith('_ms')))
print(d['total'])
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_1gpu.json 2> ${T}_bench_1gpu.err
echo "== bench: exit $?"; cut -c1-330 ${T}_bench_1gpu.json; tail -2 ${T}_bench_1gpu.err
