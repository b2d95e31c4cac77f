#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02N
timeout -s KILL 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 200 -k "frames_in_flight or fused_pipeline or graph_replay" > ${T}_tests.log 2>&1
echo "== pipeline tests: exit $?"; tail -2 ${T}_tests.log | cut -c1-200
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
