#!/bin/bash
# new default (sliding window + TMA residual + fused pool): full GPU suite, smoke, bench, layer table, launch list of one step
mkdir -p gpurun_out
T=gpurun_out/r02A
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 > ${T}_gpu_tests.log 2>&1
echo "== gpu tests: exit $?"; tail -3 ${T}_gpu_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_gpu_tests.log | head -20
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > ${T}_smoke.log 2>&1
echo "== smoke: exit $?"; tail -2 ${T}_smoke.log
timeout -s KILL 900 python bench.py --steps 40 --warmup 3 > ${T}_bench.json 2> ${T}_bench.err
echo "== bench: exit $?"; cut -c1-300 ${T}_bench.json; tail -2 ${T}_bench.err
timeout -s KILL 600 python tools/gpu_layer_table.py --mpx-only --out ${T}_layer_table.json > ${T}_layer_table.log 2>&1
echo "== layer table: exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02A_layer_table.json'))
for r in d['rows']: print(r['layer'], r['count'], round(r['mpx_ms'],3), round(r['mpx_tflops']))
print(d.get('total'))
PY
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file ${T}_launches_step.csv python tools/profile_step.py > ${T}_launches.log 2>&1
echo "== launch list: exit $?"; wc -l ${T}_launches_step.csv
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --render-size 224x224 --no-cpu-baseline --no-torch-baseline > ${T}_bench_224.json 2> ${T}_bench_224.err
echo "== bench 224: exit $?"; cut -c1-200 ${T}_bench_224.json
