#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/final2_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/final2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final2_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/final2_smoke.log
timeout 900 python bench.py > $OUT/final2_bench.json 2> $OUT/final2_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/final2_bench.json"))
print("default:", d["config"]["name"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],2), "p50", round(d["e2e"]["latency_ms_p50"],4), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), d["kernel_ms_per_step"], d["clocks"], "launches", d["gpu_launches"])
for k,v in d.get("extra",{}).items():
    print("  extra", k, "value", round(v["value"],1), "ms/step", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), "p50", round(v["e2e"]["latency_ms_p50"],4), v["kernel_ms_per_step"])
PY
RAGERA_FUSE_K5=0 timeout 300 python bench.py --workload c1 --no-extra > $OUT/c1_nofuse.json 2> $OUT/c1_nofuse.err
python - <<PY
import json
v=json.load(open("$OUT/c1_nofuse.json")); print("c1 RAGERA_FUSE_K5=0: value", round(v["value"],1), "ms/step", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), v["kernel_ms_per_step"])
PY
