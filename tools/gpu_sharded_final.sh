#!/bin/bash
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR tests/tools/sharded_check.py > $OUT/sharded_check_final_n$N.log 2>&1; echo "sharded_check n=$N exit $?"
grep -i "parity\|MISMATCH\|error" $OUT/sharded_check_final_n$N.log | head -5
