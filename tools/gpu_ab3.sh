#!/bin/bash
# A/B of one env knob, no tests: tools/gpu_ab3.sh VAR "v1 v2" "workloads" [steps]
OUT=gpurun_out; mkdir -p $OUT
VAR=$1; VALS=$2; WLS=${3:-"c5"}; STEPS=${4:-12}
for V in $VALS; do
  for W in $WLS; do
    env $VAR=$V timeout 400 python bench.py --workload $W --no-extra --steps $STEPS --warmup 3 > $OUT/${W}_${VAR}_$V.json 2> $OUT/${W}_${VAR}_$V.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/${W}_${VAR}_$V.json")); print("$W $VAR=$V: value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["certified"]["last_step"])
except Exception as e: print("$W $VAR=$V parse failed", e)
PY
  done
done
