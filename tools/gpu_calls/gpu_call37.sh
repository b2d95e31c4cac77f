#!/bin/bash
# layer3/4 pair kernel: staged half-tile epilogue (mode bit 25) -- parity, then per-layer and bench A/B
mkdir -p gpurun_out
T=gpurun_out/r02I
timeout -s KILL 420 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 150 -k "cta_pair_kernel" > ${T}_tests.log 2>&1
rc=$?
echo "== tests: exit $rc"; tail -3 ${T}_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests.log | head -20
if [ $rc -ne 0 ]; then exit 0; fi
MPX_CONV_MODE=60866571 timeout -s KILL 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_pipeline.py tests/test_zz_gpu_fullsize.py -m gpu -q --timeout 300 -k "resnet34_engine or wide_resnet or network_with or pipeline_matches or fused_pipeline_equals or full_size" > ${T}_tests2.log 2>&1
echo "== network / pipeline / full-size tests under the staged mode: exit $?"; tail -3 ${T}_tests2.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests2.log | head -20
timeout -s KILL 600 python tools/gpu_layer_table.py --mpx-only --conv-modes 60866571 --out ${T}_layer_table.json > ${T}_layer_table.log 2>&1
echo "== layer table: exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02I_layer_table.json'))
for r in d['rows']:
    if 'layer3' in r['layer'] or 'layer4' in r['layer']: print(r['layer'], r['count'], round(r['mpx_ms'],3), round(r['mpx_mode60866571_ms'],3))
print(d.get('total'))
PY
for M in 27312139 60866571 27312139 60866571; do
MPX_CONV_MODE=$M timeout -s KILL 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_$M.json 2> ${T}_bench_$M.err
echo "== bench mode $M: exit $?"; python - <<PY
import json
d=json.loads(open("${T}_bench_$M.json").read().splitlines()[-1])
print(round(d["ms_per_step"],3), "single", round(d["single_frame"]["ms_per_step"],3), "conv_ms", round(d["roofline"]["conv_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3))
PY
done
