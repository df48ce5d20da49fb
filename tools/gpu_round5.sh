#!/bin/bash
# full GPU suite + K2 prefetch sweep + C4 + mode-2 cycle breakdown
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r5_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r5_pytest.log
for PF in 0 4 16 32; do
  RAGERA_K2_PREFETCH=$PF timeout 200 python bench.py --workload c2b --no-extra --steps 30 --warmup 5 > $OUT/c2b_pf$PF.json 2> $OUT/c2b_pf$PF.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/c2b_pf$PF.json")); print("prefetch $PF: value", round(d["value"]), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), d["clocks"])
except Exception as e: print("pf $PF parse failed", e)
PY
done
timeout 300 python bench.py --workload c4 --no-extra --steps 20 --warmup 5 > $OUT/c4_r5.json 2> $OUT/c4_r5.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/c4_r5.json")); print("c4: value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), d["kernel_ms_per_step"], d["certified"], "e2e", round(d["e2e"]["value"]))
except Exception as e: print("c4 parse failed", e)
PY
RAGERA_K2_MODE=2 RAGERA_K2_PROF=1 timeout 200 python bench.py --workload c2b --no-extra --steps 20 --warmup 5 > $OUT/c2b_m2prof.json 2> $OUT/c2b_m2prof.err
grep -A 3 "k2 pair prof" $OUT/c2b_m2prof.err | head -4
