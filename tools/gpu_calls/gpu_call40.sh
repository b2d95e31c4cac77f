#!/bin/bash
# compacted pooled epilogue: parity, stem timing, bench
mkdir -p gpurun_out
T=gpurun_out/r02K
timeout -s KILL 420 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout 150 -k "fused_maxpool or network_with_fused or stem_space or pair_window64 or resnet34_engine" > ${T}_tests.log 2>&1
rc=$?
echo "== tests: exit $rc"; tail -3 ${T}_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_tests.log | head -20
if [ $rc -ne 0 ]; then exit 0; fi
timeout -s KILL 300 python tools/gpu_pool_ab.py 576 > ${T}_pool_ab.json 2> ${T}_pool_ab.err
echo "== pool ab: exit $?"; cat ${T}_pool_ab.json; tail -2 ${T}_pool_ab.err
for i in 1 2; do
timeout -s KILL 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_$i.json 2> ${T}_bench_$i.err
echo "== bench $i: exit $?"; python - <<PY
import json
d=json.loads(open("${T}_bench_$i.json").read().splitlines()[-1])
print(round(d["ms_per_step"],3), "single", round(d["single_frame"]["ms_per_step"],3), "conv_ms", round(d["roofline"]["conv_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3))
PY
done
