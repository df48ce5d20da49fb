#!/bin/bash
# quick K2 iteration: tensor parity tests, c2b bench (mode 0 and 2), per-warp cycle breakdown
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-x}
timeout 400 python -m pytest tests/test_gpu_tensor.py -x -q > $OUT/k2_pytest_$TAG.log 2>&1; echo "k2 pytest exit $?"; tail -25 $OUT/k2_pytest_$TAG.log
for MODE in 0 2; do
  RAGERA_K2_MODE=$MODE timeout 200 python bench.py --workload c2b --no-extra --steps 30 --warmup 5 > $OUT/c2b_${TAG}_m$MODE.json 2> $OUT/c2b_${TAG}_m$MODE.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/c2b_${TAG}_m$MODE.json")); print("mode $MODE: value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), d["kernel_ms_per_step"], d["certified"], d["clocks"])
except Exception as e: print("mode $MODE parse failed", e)
PY
done
RAGERA_K2_PROF=1 timeout 200 python bench.py --workload c2b --no-extra --steps 20 --warmup 5 > $OUT/c2b_${TAG}_prof.json 2> $OUT/c2b_${TAG}_prof.err
grep -A 8 "k2 pair prof" $OUT/c2b_${TAG}_prof.err | head -9
