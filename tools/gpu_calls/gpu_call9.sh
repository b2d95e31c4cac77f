#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02i
timeout 900 python -m pytest tests -m gpu -q -rs > ${T}_gputests.log 2>&1
echo "== gpu tests: exit $?"; grep -E "passed|failed|error" ${T}_gputests.log | tail -3; grep -E "^E  |FAILED|object [0-9]" ${T}_gputests.log | head -20
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > ${T}_bench_1gpu.json 2> ${T}_bench_1gpu.err
echo "== bench: exit $?"; cut -c1-330 ${T}_bench_1gpu.json; tail -2 ${T}_bench_1gpu.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file ${T}_launches_step.csv python tools/profile_step.py > ${T}_ncu_step.log 2>&1
echo "== ncu launch list: exit $?"; python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02i_launches_step.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
t=collections.defaultdict(float); c=collections.Counter()
for r in rows[1:]:
    if r[mi]=='gpu__time_duration.sum':
        name=r[ki].split('(')[0][:60]; t[name]+=float(r[vi].replace(',','')); c[name]+=1
tot=sum(t.values())
for k,v in sorted(t.items(), key=lambda kv:-kv[1])[:16]: print(f"{v/1e6:8.3f} ms {100*v/tot:5.1f}% x{c[k]:4d} {k}")
print('total ms', tot/1e6)
PY
