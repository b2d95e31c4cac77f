#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02g
timeout 900 python -m pytest tests -m gpu -q -rs > ${T}_gputests.log 2>&1
echo "== gpu tests: exit $?"; grep -E "passed|failed|error" ${T}_gputests.log | tail -3; grep -E "^E  |FAILED" ${T}_gputests.log | head -20
timeout 600 python tools/gpu_layer_table.py --mpx-only --conv-modes 11,139 --out ${T}_layer_modes.json > ${T}_layer_modes.log 2>&1
echo "== layer modes: exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_layer_modes.json'))
for r in d['rows']:
    print(r['layer'], r['count'], ' '.join(f"{k[4:-3]}={v:.3f}" for k,v in r.items() if k.startswith('mpx_') and k.endswith('_ms')))
print(d['total'])
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_1gpu.json 2> ${T}_bench_1gpu.err
echo "== bench: exit $?"; cut -c1-330 ${T}_bench_1gpu.json; tail -2 ${T}_bench_1gpu.err
