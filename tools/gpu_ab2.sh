#!/bin/bash
# A/B of one env knob on a list of workloads: tools/gpu_ab2.sh VAR "v1 v2" "c2b c5"
OUT=gpurun_out; mkdir -p $OUT
VAR=$1; VALS=$2; WLS=${3:-"c2b c5"}
for V in $VALS; do
  for W in $WLS; do
    env $VAR=$V timeout 400 python bench.py --workload $W --no-extra --steps 12 --warmup 3 > $OUT/${W}_${VAR}_$V.json 2> $OUT/${W}_${VAR}_$V.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/${W}_${VAR}_$V.json")); print("$W $VAR=$V: value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), d["clocks"]["reasons"], d["certified"]["last_step"])
except Exception as e: print("$W $VAR=$V parse failed", e)
PY
  done
done
timeout 300 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -2
