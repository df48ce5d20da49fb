#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for L in 2 0; do
  for W in c2b c5; do
    RAGERA_K2_LOCKSTEP=$L timeout 400 python bench.py --workload $W --no-extra --steps 12 --warmup 3 > $OUT/${W}_lock$L.json 2> $OUT/${W}_lock$L.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/${W}_lock$L.json")); print("$W lockstep=$L: value", round(d["value"]), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), d["clocks"]["reasons"], d["certified"]["last_step"])
except Exception as e: print("$W lockstep=$L parse failed", e)
PY
  done
done
timeout 300 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -2
