#!/bin/bash
# First GPU call of a round: the whole -m gpu suite, the bench line (with the torch / cuDNN bar and the CPU baseline), the
# per-layer table against cuDNN, then everything that was written without a GPU at hand (the experimental kernel
# variants, each group in its own process: bounded mbarrier waits trap after ~4 s and poison that process only) and an A/B
# of the variants whose parity tests passed.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_first_call.sh'
mkdir -p gpurun_out
T=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${T}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rs --durations=12 > ${T}_gputests.log 2>&1
echo "== gpu tests: exit $?"; grep -E "passed|failed|error" ${T}_gputests.log | tail -3
grep -E "^\[|FAILED|Error|assert" ${T}_gputests.log | head -60
timeout 600 python bench.py --steps 20 --warmup 3 > ${T}_bench_1gpu.json 2> ${T}_bench_1gpu.err
echo "== bench: exit $?"; cut -c1-1500 ${T}_bench_1gpu.json; tail -3 ${T}_bench_1gpu.err
timeout 300 python bench.py --steps 20 --warmup 3 --render-size 224x224 --no-cpu-baseline --no-torch-baseline > ${T}_bench_1gpu_224.json 2> ${T}_bench_1gpu_224.err
echo "== bench 224: exit $?"; cut -c1-400 ${T}_bench_1gpu_224.json
timeout 600 python tools/gpu_layer_table.py --out ${T}_layer_table.json > ${T}_layer_table.log 2>&1
echo "== layer table: exit $?"; tail -2 ${T}_layer_table.log | cut -c1-1200

export MPX_EXPERIMENTAL=1
PASSED="11"
run() {  # name, mode value for the A/B, pytest -k expression
  local name=$1 mode=$2; shift 2
  timeout 300 python -m pytest tests/test_gpu_net.py -q -x -k "$1" > "${T}_exp_${name}.log" 2>&1
  local rc=$?
  echo "== experimental ${name}: exit ${rc}  $(tail -1 ${T}_exp_${name}.log)"
  if [ ${rc} -eq 0 ] && [ -n "${mode}" ]; then PASSED="${PASSED},${mode}"; fi
}
run observers       2059  "experimental_window_observers"
run pairwin_l34     4107  "experimental_pair_window_kernel and (l3 or l4 or odd_size) and not l2"
run pairwin_l2_128  8203  "experimental_pair_window_kernel and l2_128wide"
run pairs_l2        16395 "experimental_pair_window_kernel and l2_pairs"
run pairs_64        32779 "experimental_pair_window64"
unset MPX_EXPERIMENTAL
# 11 default | +2048 gated refills | +4096 layer3/4 pair-window | +8192 layer2 128-wide pair-window | +16384 layer2 pairs |
# +32768 stem+layer1 pairs; the default brackets the list so that drift shows
timeout 900 python tools/gpu_ab.py --conv "${PASSED},11" --steps 20 --rounds 2 --timeout 120 --out ${T}_ab_modes.json > ${T}_ab_modes.log 2>&1
echo "== A/B (${PASSED},11): exit $?"; tail -70 ${T}_ab_modes.log
