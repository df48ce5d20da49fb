#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { # name, env...
  name=$1; shift
  env "$@" timeout 100 python bench.py --workload c4 --no-extra --steps 20 --warmup 4 > $OUT/c4_$name.json 2> $OUT/c4_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/c4_$name.json")); print("$name", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), d["kernel_ms_per_step"], "roof", round(d["roofline"]["achieved"],1))
except Exception as e: print("$name failed", e)
PY
}
run base RAGERA_X=0
run pf0 RAGERA_K2_PREFETCH=0
run pf24 RAGERA_K2_PREFETCH=24
run prof RAGERA_K2_PROF=1
grep -v "^$" $OUT/c4_prof.err | tail -40
