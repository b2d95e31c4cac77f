#!/bin/bash
# sliding-window pair kernel (mode bit 23): parity first (bounded: a barrier bug would hang), then per-layer and bench A/B
mkdir -p gpurun_out
T=gpurun_out/r02y
timeout -s KILL 420 python -m pytest tests/test_gpu_net.py -m gpu -q -x --timeout 150 -k "pair_window64 or fused_maxpool" > ${T}_tests.log 2>&1
rc=$?
echo "== sliding tests: exit $rc"; tail -3 ${T}_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests.log | head -20
if [ $rc -ne 0 ]; then exit 0; fi
MPX_CONV_MODE=10534923 timeout -s KILL 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_pipeline.py -m gpu -q -x --timeout 200 -k "resnet34_engine or network_with or small_batch or pipeline_matches or fused_pipeline_equals" > ${T}_tests2.log 2>&1
echo "== network / pipeline tests under the sliding mode: exit $?"; tail -3 ${T}_tests2.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests2.log | head -20
timeout -s KILL 600 python tools/gpu_layer_table.py --mpx-only --conv-modes 10534923 --out ${T}_layer_table.json > ${T}_layer_table.log 2>&1
echo "== layer table: exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02y_layer_table.json'))
for r in d['rows'][:4]: print(r['layer'], r['count'], round(r['mpx_ms'],3), round(r['mpx_mode10534923_ms'],3))
print(d.get('total'))
PY
for M in 2146315 10534923 2146315 10534923; do
MPX_CONV_MODE=$M timeout -s KILL 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_$M.json 2> ${T}_bench_$M.err
echo "== bench mode $M: exit $?"; python - <<PY
import json
d=json.loads(open("${T}_bench_$M.json").read().splitlines()[-1])
print(round(d["ms_per_step"],3), "single", round(d["single_frame"]["ms_per_step"],3), "conv_ms", round(d["roofline"]["conv_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3))
PY
done
