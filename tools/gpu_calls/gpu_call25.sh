#!/bin/bash
# full GPU suite on the new default (fused max-pool), then late-PDL A/B with longer runs
mkdir -p gpurun_out
T=gpurun_out/r02x
timeout 1500 python -m pytest tests -m gpu -q > ${T}_gpu_tests.log 2>&1
echo "== gpu tests: exit $?"; tail -3 ${T}_gpu_tests.log | cut -c1-300; grep -E "^(E |FAILED)" ${T}_gpu_tests.log | head -20
run_bench() {  # name, conv mode, extra args
  MPX_CONV_MODE=$2 timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-torch-baseline $3 > ${T}_bench_$1.json 2> ${T}_bench_$1.err
  echo "== bench $1 (mode $2 $3): exit $?"; python - <<PY
import json
d=json.loads(open("${T}_bench_$1.json").read().splitlines()[-1])
print(round(d["ms_per_step"],3), "single", round(d["single_frame"]["ms_per_step"],3) if d.get("single_frame") else None, "e2e", round(d["e2e"]["value"]), "conv_ms", round(d["roofline"]["conv_ms_per_step"],3))
PY
}
run_bench pool_a 2146315 ""
run_bench poolpdl_a 6340619 ""
run_bench pool_b 2146315 ""
run_bench poolpdl_b 6340619 ""
run_bench pool_fif3 2146315 "--frames-in-flight 3"
run_bench poolpdl_fif3 6340619 "--frames-in-flight 3"
run_bench poolpdl_nosplitcap 6340619 ""
MPX_CONV_MODE=6602763 run_bench poolpdl_splitcap2 6602763 ""
