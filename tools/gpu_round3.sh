#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(d["config"]["name"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "p50", round(d["e2e"]["latency_ms_p50"],4), "roof", d["roofline"]["bound"], round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), d["kernel_ms_per_step"], d.get("certified"), d["e2e"].get("uncertified_after_escalation"), d["clocks"])
except Exception as e: print("parse failed", sys.argv[1], e)
PY
}
timeout 400 python bench.py --workload c4 --no-extra --steps 40 --warmup 5 > $OUT/c4_bench.json 2> $OUT/c4_bench.err; echo "c4 exit $?"; show $OUT/c4_bench.json; tail -2 $OUT/c4_bench.err
timeout 600 python bench.py --workload c5 --no-extra --steps 5 --warmup 3 > $OUT/c5_n1_bench.json 2> $OUT/c5_n1_bench.err; echo "c5 exit $?"; show $OUT/c5_n1_bench.json; tail -2 $OUT/c5_n1_bench.err
CMD="python bench.py --workload c2b --no-extra --steps 4 --warmup 3"
timeout 300 $CMD > $OUT/plain_k2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_pair -s 3 -c 1 -o $OUT/r01_k2_pair_full $CMD > $OUT/ncu_k2_full.log 2>&1
echo "ncu k2 exit $?"; tail -2 $OUT/ncu_k2_full.log
CMD3="python bench.py --workload c3 --no-extra --steps 3 --warmup 3"
timeout 300 $CMD3 > $OUT/plain_c3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:k1_stream -s 3 -c 1 -o $OUT/r01_k1_c3_full $CMD3 > $OUT/ncu_k1c3_full.log 2>&1
echo "ncu k1 c3 exit $?"; tail -2 $OUT/ncu_k1c3_full.log
