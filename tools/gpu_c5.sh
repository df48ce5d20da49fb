#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_c5.sh N'
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 800 $TR bench.py --gpus $N --workload c5 --no-extra --steps 20 --warmup 4 > $OUT/bench_c5_p2p_n$N.json 2> $OUT/bench_c5_p2p_n$N.err; echo "bench c5 n=$N exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_c5_p2p_n$N.json")); print(d["config"]["name"], "n", d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["kernel_ms_per_step"], d.get("certified"), d["dtype"]); print(d.get("per_rank_kernel_ms"))
PY
grep -i "error\|trap" $OUT/bench_c5_p2p_n$N.err | head -3
