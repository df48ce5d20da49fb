#!/bin/bash
OUT=gpurun_out
for S in 6 14 22; do
  RAGERA_BENCH_SLACK=$S timeout 200 python bench.py --workload c2b --no-extra --steps 30 --warmup 5 > $OUT/c2b_s$S.json 2> $OUT/c2b_s$S.err
  python - <<PY
import json
d=json.load(open("$OUT/c2b_s$S.json")); print("slack $S: value", round(d["value"]), "k2 ms", round(d["roofline"]["avg_launch_ms"],3), "TF", round(d["roofline"]["achieved"],1), "e2e", round(d["e2e"]["value"]), d["certified"], "uncert-after-esc", d["e2e"]["uncertified_after_escalation"], "p50", round(d["e2e"]["latency_ms_p50"],3))
PY
done
