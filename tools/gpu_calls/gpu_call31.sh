#!/bin/bash
# heads gated across the two frames in flight (FramePipeline.serialize_heads) vs overlapping heads; layer2 L2-prefetch A/B
mkdir -p gpurun_out
T=gpurun_out/r02C
timeout -s KILL 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_net.py -m gpu -q -x --timeout 300 -k "frames_in_flight or layer2_pair_window or window_and_im2col" > ${T}_tests.log 2>&1
echo "== tests: exit $?"; tail -3 ${T}_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests.log | head
run_bench() {
  timeout -s KILL 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-torch-baseline $2 > ${T}_bench_$1.json 2> ${T}_bench_$1.err
  echo "== bench $1 ($2): exit $?"; python - <<PY
import json
d=json.loads(open("${T}_bench_$1.json").read().splitlines()[-1])
print(round(d["ms_per_step"],3), "hyp/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["single_frame"]["ms_per_step"],3) if d.get("single_frame") else None, "conv_ms", round(d["roofline"]["conv_ms_per_step"],3))
PY
}
run_bench gated_a ""
run_bench overlap_a "--overlap-heads"
run_bench gated_b ""
run_bench overlap_b "--overlap-heads"
run_bench gated_fif3 "--frames-in-flight 3"
timeout -s KILL 600 python tools/gpu_layer_table.py --mpx-only --out ${T}_layer_table.json > ${T}_layer_table.log 2>&1
echo "== layer table: exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02C_layer_table.json'))
for r in d['rows']:
    if 'layer2' in r['layer'] or 'layer1' in r['layer']: print(r['layer'], r['count'], round(r['mpx_ms'],3))
print(d.get('total'))
PY
