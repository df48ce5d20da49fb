#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_multi2.sh N'  — round-1 final multi-GPU numbers (p2p exchange; NCCL beside it for C3)
N=${1:-8}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(d["config"]["name"], "n_gpus", d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "p50", round(d["e2e"]["latency_ms_p50"],4), "roof", d["roofline"]["bound"], round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), d["kernel_ms_per_step"], d.get("certified"), d["clocks"]["reasons"])
except Exception as e: print("parse failed", sys.argv[1], e)
PY
}
timeout 300 $TR tests/tools/sharded_check.py > $OUT/sharded_check_p2p_n$N.log 2>&1; echo "sharded check exit $?"; grep -E "parity|MISMATCH" $OUT/sharded_check_p2p_n$N.log | head -3
timeout 400 $TR bench.py --gpus $N --no-extra > $OUT/bench_c3_p2p_n$N.json 2> $OUT/bench_c3_p2p_n$N.err; echo "bench c3 p2p n=$N exit $?"; show $OUT/bench_c3_p2p_n$N.json
RAGERA_COMM=nccl timeout 400 $TR bench.py --gpus $N --no-extra > $OUT/bench_c3_nccl_n$N.json 2> $OUT/bench_c3_nccl_n$N.err; echo "bench c3 nccl n=$N exit $?"; show $OUT/bench_c3_nccl_n$N.json
timeout 500 $TR bench.py --gpus $N --workload c5 --no-extra --steps 30 --warmup 4 > $OUT/bench_c5_p2p_n$N.json 2> $OUT/bench_c5_p2p_n$N.err; echo "bench c5 n=$N exit $?"; show $OUT/bench_c5_p2p_n$N.json
timeout 300 $TR bench.py --gpus $N --workload c2b --no-extra --steps 50 --warmup 5 > $OUT/bench_c2b_p2p_n$N.json 2> $OUT/bench_c2b_p2p_n$N.err; echo "bench c2b n=$N exit $?"; show $OUT/bench_c2b_p2p_n$N.json
