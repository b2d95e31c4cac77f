#!/bin/bash
# second GPU call of round 2: tiled rasteriser parity + microbench, per-layer timing of the pair-kernel bits, ncu captures
mkdir -p gpurun_out
T=gpurun_out/r02b
timeout 900 python -m pytest tests -m gpu -q -rs -x --deselect tests/test_zz_gpu_fullsize.py > ${T}_gputests.log 2>&1
echo "== gpu tests (without full size): exit $?"; grep -E "passed|failed|error" ${T}_gputests.log | tail -3; grep -E "^\[|^E  |FAILED" ${T}_gputests.log | head -40
timeout 900 python -m pytest tests/test_zz_gpu_fullsize.py -m gpu -q -rs -s > ${T}_fullsize.log 2>&1
echo "== full size: exit $?"; grep -E "^\[|passed|failed|^E  " ${T}_fullsize.log | head -80
timeout 300 python tools/bench_raster.py > ${T}_raster_microbench.json 2> ${T}_raster_microbench.err
echo "== raster microbench: exit $?"; cat ${T}_raster_microbench.json; tail -3 ${T}_raster_microbench.err
timeout 600 python tools/gpu_layer_table.py --mpx-only --conv-modes 16395,32779,49163,81931,114699 --out ${T}_layer_modes.json > ${T}_layer_modes.log 2>&1
echo "== layer modes: exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_layer_modes.json'))
for r in d['rows']:
    print(r['layer'], r['count'], ' '.join(f"{k[4:-3]}={v:.3f}" for k,v in r.items() if k.startswith('mpx_') and k.endswith('_ms')))
print(d['total'])
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_1gpu.json 2> ${T}_bench_1gpu.err
echo "== bench: exit $?"; cut -c1-330 ${T}_bench_1gpu.json
# launch list of one step, and a full capture of the rasteriser of the 576-hypothesis coarse stage
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file ${T}_launches_step.csv python tools/profile_step.py > ${T}_ncu_step.log 2>&1
echo "== ncu launch list: exit $?"; python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02b_launches_step.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
t=collections.defaultdict(float); c=collections.Counter()
for r in rows[1:]:
    if r[mi]=='gpu__time_duration.sum':
        name=r[ki].split('(')[0][:60]; t[name]+=float(r[vi].replace(',','')); c[name]+=1
tot=sum(t.values())
for k,v in sorted(t.items(), key=lambda kv:-kv[1])[:14]: print(f"{v/1e6:8.3f} ms {100*v/tot:5.1f}% x{c[k]:4d} {k}")
print('total ms', tot/1e6)
PY
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:raster_tiled -o ${T}_ncu_raster_tiled python tools/profile_step.py --stage coarse > ${T}_ncu_raster.log 2>&1
echo "== ncu raster: exit $?"; ls -la ${T}_ncu_raster_tiled.ncu-rep 2>/dev/null
