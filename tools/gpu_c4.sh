#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r5_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r5_pytest.log
for W in c4 c4f; do
timeout 400 python bench.py --workload $W --no-extra --steps 40 --warmup 5 > $OUT/${W}_bench.json 2> $OUT/${W}_bench.err; echo "$W exit $?"
python - $OUT/${W}_bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(d["config"]["name"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "roof", r["bound"], round(r["achieved"],1), "peak", r["peak"], round(r["frac"],3), r.get("hbm_achieved_GBps"), d["kernel_ms_per_step"], d.get("certified"), d["e2e"].get("uncertified_after_escalation"), d["clocks"])
except Exception as e: print("parse failed", e)
PY
done
