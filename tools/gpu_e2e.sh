#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for W in c4 c2b; do
  RAGERA_BENCH_E2E_PROF=1 timeout 300 python bench.py --workload $W --no-extra --steps 30 --warmup 5 > $OUT/${W}_e2eprof.json 2> $OUT/${W}_e2eprof.err
  python - <<PY
import json
d=json.load(open("$OUT/${W}_e2eprof.json")); print("$W: ms/step", round(d["ms_per_step"],3), d["kernel_ms_per_step"]); print("   e2e p50", round(d["e2e"]["latency_ms_p50"],3), "mean", round(1e3*d["config"]["batch"]/d["e2e"]["value"],3), d.get("e2e_kernel_ms_per_call"))
PY
done
