#!/bin/bash
# round-end verification on one B200: full GPU suite, smoke, default bench line, ncu launch list of the default workload
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/final_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/final_smoke.log
timeout 900 python bench.py > $OUT/final_bench.json 2> $OUT/final_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/final_bench.json"))
print("default:", d["config"]["name"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],2), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), d["kernel_ms_per_step"], d["clocks"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"]["value"])
for k,v in d.get("extra",{}).items():
    print("  extra", k, "value", round(v["value"],1), "ms/step", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), "roof", round(v["roofline"]["achieved"],1) if v.get("roofline") and v["roofline"].get("achieved") else None, v["kernel_ms_per_step"])
PY
CMD="python bench.py --no-extra --steps 5 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r01_launches_c3_default.csv $CMD > $OUT/ncu_launches_c3.log 2>&1
echo "ncu launches exit $?"
