#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02D
for cfg in "2 gate" "2 overlap" "1 gate"; do
timeout -s KILL 300 python tools/gpu_pipeline_timeline.py $cfg > ${T}_timeline_$(echo $cfg | tr ' ' _).json 2> ${T}_timeline.err
echo "== timeline $cfg: exit $?"; python - <<PY
import json
d=json.loads(open("${T}_timeline_$(echo $cfg | tr ' ' _).json").read().splitlines()[-1])
print("frame_ms", round(d["frame_ms"],3))
for r in d["rows"][2:10]: print(r)
PY
tail -2 ${T}_timeline.err
done
