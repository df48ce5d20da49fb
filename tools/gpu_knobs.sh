#!/bin/bash
# Knob scan: one short bench line per "name workload ENV=VAL..." spec, the scoring kernel's time printed per line.
#   gpurun -- 'bash tools/gpu_knobs.sh "c2b_default c2b" "c2b_lock0 c2b RAGERA_K2_LOCKSTEP=0" "c2b_prof c2b RAGERA_K2_PROF=1"'
# (the scans behind profiles/r02_k2_knobs.md: RAGERA_K2_STAGES / _PREFETCH / _LOCKSTEP / _MODE / _CLUSTER on c2b, c4, c5;
#  RAGERA_GRAPH=0|1 and RAGERA_SMALL_PROF=1 on c1, c2). Knobs are listed in DESIGN.md §4.
OUT=gpurun_out; mkdir -p $OUT
for spec in "$@"; do
  set -- $spec; name=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-extra ${BENCH_ARGS:---steps 20 --warmup 5} > $OUT/knob_$name.json 2> $OUT/knob_$name.err
  python - "$OUT/knob_$name.json" "$name" <<'P'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    k = j["kernel_ms_per_step"]
    print(sys.argv[2], " ".join(f"{c} {v:.4f} ms" for c, v in k.items()), f"step {j['ms_per_step']:.4f} ms", f"e2e p50 {j['e2e']['latency_ms_p50']:.4f} ms",
          "certified", j["certified"]["timed_steps"], "/", j["certified"]["timed_steps_of"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
P
  grep -A8 "prof\]" $OUT/knob_$name.err | head -12
done
