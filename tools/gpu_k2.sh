#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_tensor.py -x -q > $OUT/k2_pytest.log 2>&1; echo "k2 (pair) pytest exit $?" | tee -a $OUT/k2_pytest.log
tail -15 $OUT/k2_pytest.log
RAGERA_K2_IMPL=1 timeout 240 python -m pytest tests/test_gpu_tensor.py -x -q > $OUT/k2s_pytest.log 2>&1; echo "k2 (single) pytest exit $?" | tee -a $OUT/k2s_pytest.log
tail -4 $OUT/k2s_pytest.log
for MODE in 0 1 2; do
  RAGERA_K2_MODE=$MODE timeout 200 python bench.py --workload c2b --no-extra --steps 30 --warmup 5 > $OUT/c2b_mode$MODE.json 2> $OUT/c2b_mode$MODE.err
  echo "mode $MODE exit $?"; python - <<PY
import json
try:
    d=json.load(open("$OUT/c2b_mode$MODE.json")); print("mode $MODE: value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), d["kernel_ms_per_step"], d["certified"])
except Exception as e: print("mode $MODE parse failed", e)
PY
done
