#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/r2_pytest.log; tail -5 $OUT/r2_pytest.log
timeout 600 python bench.py > $OUT/r2_bench.json 2> $OUT/r2_bench.err; echo "bench exit $?"; tail -3 $OUT/r2_bench.err
python - <<PY
import json
d=json.load(open("$OUT/r2_bench.json"))
def show(name,x): print(name, "value", round(x["value"],1), "ms/step", round(x["ms_per_step"],4), "e2e", round(x["e2e"]["value"],1), "p50", round(x["e2e"]["latency_ms_p50"],4), "roof", x["roofline"]["bound"], round(x["roofline"]["achieved"],1), round(x["roofline"]["frac"],3), x["kernel_ms_per_step"], x.get("certified"), x["e2e"].get("uncertified_after_escalation"))
show("c3", d)
for k,v in d.get("extra",{}).items(): show(k, v)
print("cpu", d.get("cpu_baseline"))
PY
