#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 230 python bench.py > $OUT/bench_v7.json 2> $OUT/bench_v7.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_v7.json"))
print("default:", d["config"]["name"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],2), "p50", round(d["e2e"]["latency_ms_p50"],4), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), d["kernel_ms_per_step"], d["clocks"], "launches", d["gpu_launches"], "wall", d.get("wall_s"))
for k,v in d.get("extra",{}).items():
    print("  extra", k, "value", round(v["value"],1), "ms/step", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), "p50", round(v["e2e"]["latency_ms_p50"],4), v["kernel_ms_per_step"])
PY
