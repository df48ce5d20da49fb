#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/sharded_check.py > $OUT/sharded_check_n$N.log 2>&1; echo "sharded check exit $?"; tail -5 $OUT/sharded_check_n$N.log
timeout 600 $TR bench.py --gpus $N > $OUT/bench_c3_n$N.json 2> $OUT/bench_c3_n$N.err; echo "bench c3 n=$N exit $?"; tail -c 1800 $OUT/bench_c3_n$N.json; tail -3 $OUT/bench_c3_n$N.err
timeout 600 $TR bench.py --gpus $N --workload c2b --steps 50 --warmup 5 > $OUT/bench_c2b_n$N.json 2> $OUT/bench_c2b_n$N.err; echo "bench c2b n=$N exit $?"; tail -c 1500 $OUT/bench_c2b_n$N.json; tail -3 $OUT/bench_c2b_n$N.err
