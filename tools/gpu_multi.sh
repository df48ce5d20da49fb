#!/bin/bash
# multi-GPU validation: usage  gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/tools/sharded_check.py > $OUT/r2_sharded_check_n$N.log 2>&1; echo "sharded_check N=$N exit $?" | tee -a $OUT/r2_sharded_check_n$N.log
tail -3 $OUT/r2_sharded_check_n$N.log
RAGERA_COMM=nccl timeout 900 $TR --master-port 29512 tests/tools/sharded_check.py > $OUT/r2_sharded_check_n${N}_nccl.log 2>&1; echo "sharded_check (nccl) N=$N exit $?" | tee -a $OUT/r2_sharded_check_n${N}_nccl.log
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 100 --warmup 10 > $OUT/r2_bench_c3_n$N.json 2> $OUT/r2_bench_c3_n$N.err; echo "bench c3 N=$N exit $?"
tail -c 600 $OUT/r2_bench_c3_n$N.err
