#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_parity.py -x -q > $OUT/r7_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r7_pytest.log
for W in c2b c4 c4f; do
  timeout 300 python bench.py --workload $W --no-extra --steps 30 --warmup 5 > $OUT/${W}_r7.json 2> $OUT/${W}_r7.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${W}_r7.json")); print("$W: value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), d["kernel_ms_per_step"], d["clocks"]["reasons"], "e2e", round(d["e2e"]["value"]), d["certified"])
except Exception as e: print("$W parse failed", e)
PY
done
