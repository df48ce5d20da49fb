#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/k4_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/k4_pytest.log
timeout 300 python bench.py --workload c2b --no-extra --steps 30 --warmup 5 > $OUT/c2b_k4.json 2> $OUT/c2b_k4.err
python - <<PY
import json
d=json.load(open("$OUT/c2b_k4.json")); print("c2b: value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), d["kernel_ms_per_step"], "e2e", round(d["e2e"]["value"]))
PY
