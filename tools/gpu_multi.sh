#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(d["config"]["name"], "n_gpus", d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "p50", round(d["e2e"]["latency_ms_p50"],4), "roof", d["roofline"]["bound"], round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), d["kernel_ms_per_step"], d.get("certified"), d["clocks"])
except Exception as e: print("parse failed", sys.argv[1], e)
PY
}
timeout 600 $TR tests/tools/sharded_check.py > $OUT/sharded_check_n$N.log 2>&1; echo "sharded check exit $?"; grep -E "parity|MISMATCH" $OUT/sharded_check_n$N.log | head
timeout 600 $TR bench.py --gpus $N > $OUT/bench_c3_n$N.json 2> $OUT/bench_c3_n$N.err; echo "bench c3 n=$N exit $?"; show $OUT/bench_c3_n$N.json; grep -v "OMP_NUM\|\*\*\*" $OUT/bench_c3_n$N.err | tail -3
timeout 600 $TR bench.py --gpus $N --workload c5 --steps 30 --warmup 4 > $OUT/bench_c5_n$N.json 2> $OUT/bench_c5_n$N.err; echo "bench c5 n=$N exit $?"; show $OUT/bench_c5_n$N.json; grep -v "OMP_NUM\|\*\*\*" $OUT/bench_c5_n$N.err | tail -3
timeout 600 $TR bench.py --gpus $N --workload c2b --steps 50 --warmup 5 > $OUT/bench_c2b_n$N.json 2> $OUT/bench_c2b_n$N.err; echo "bench c2b n=$N exit $?"; show $OUT/bench_c2b_n$N.json
