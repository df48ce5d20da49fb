#!/bin/bash
# A/B of one K2 env knob on c2b (and c4): tools/gpu_k2ab.sh VAR v1 v2 ...
OUT=gpurun_out; mkdir -p $OUT
VAR=$1; shift
timeout 400 python -m pytest tests/test_gpu_tensor.py -x -q > $OUT/k2_pytest_ab.log 2>&1; echo "k2 pytest exit $?"; tail -3 $OUT/k2_pytest_ab.log
for V in "$@"; do
  for W in c2b c4; do
    env $VAR=$V timeout 300 python bench.py --workload $W --no-extra --steps 30 --warmup 5 > $OUT/${W}_${VAR}_$V.json 2> $OUT/${W}_${VAR}_$V.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/${W}_${VAR}_$V.json")); print("$W $VAR=$V: value", round(d["value"]), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), d["clocks"]["reasons"], d["certified"]["last_step"])
except Exception as e: print("$W $VAR=$V parse failed", e)
PY
  done
  env $VAR=$V RAGERA_K2_PROF=1 timeout 200 python bench.py --workload c2b --no-extra --steps 20 --warmup 5 > $OUT/c2b_${VAR}_${V}_prof.json 2> $OUT/c2b_${VAR}_${V}_prof.err
  grep -A 4 "k2 pair prof" $OUT/c2b_${VAR}_${V}_prof.err | head -5 | cut -c1-220
done
