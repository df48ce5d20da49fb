#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_p2p.sh N'   — peer-to-peer exchange vs the NCCL fallback
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(d["config"]["name"], "n_gpus", d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "p50", round(d["e2e"]["latency_ms_p50"],4), "roof", round(d["roofline"]["achieved"],1), d["kernel_ms_per_step"], d.get("certified"))
except Exception as e: print("parse failed", sys.argv[1], e)
PY
}
timeout 300 $TR tests/tools/sharded_check.py > $OUT/sharded_check_p2p_n$N.log 2>&1; echo "sharded check (p2p) exit $?"; grep -E "parity|MISMATCH|rror" $OUT/sharded_check_p2p_n$N.log | head
for MODE in p2p nccl; do
  RAGERA_COMM=$MODE timeout 400 $TR bench.py --gpus $N --workload c2 --no-extra --steps 300 --warmup 20 > $OUT/bench_c2_${MODE}_n$N.json 2> $OUT/bench_c2_${MODE}_n$N.err; echo "bench c2 $MODE n=$N exit $?"; show $OUT/bench_c2_${MODE}_n$N.json; grep -i "error\|trap\|fail" $OUT/bench_c2_${MODE}_n$N.err | head -3
done
RAGERA_COMM=p2p timeout 400 $TR bench.py --gpus $N --workload c2b --no-extra --steps 30 --warmup 5 > $OUT/bench_c2b_p2p_n$N.json 2> $OUT/bench_c2b_p2p_n$N.err; echo "bench c2b p2p n=$N exit $?"; show $OUT/bench_c2b_p2p_n$N.json
