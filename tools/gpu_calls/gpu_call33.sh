#!/bin/bash
# SMs reserved for the tail x split-K cap of the tail's convolutions (two frames in flight, heads gated)
mkdir -p gpurun_out
T=gpurun_out/r02E
for R in 0 4 8 12 16; do for M in 10534923 10797067 11059211; do
MPX_CONV_MODE=$M timeout -s KILL 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-torch-baseline --reserve-sms $R > ${T}_bench_r${R}_m${M}.json 2> ${T}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("${T}_bench_r${R}_m${M}.json").read().splitlines()[-1])
    print("reserve $R mode $M:", round(d["ms_per_step"],3), "ms  hyp/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["single_frame"]["ms_per_step"],3))
except Exception as e:
    print("reserve $R mode $M: failed", e)
PY
done; done
tail -3 ${T}_bench.err
