#!/bin/bash
# ncu evidence for the round-1 final K2: full capture of one launch + launch list of a batch-1024 step
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload c2b --no-extra --steps 4 --warmup 3"
timeout 300 $CMD > $OUT/plain_k2_v4.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_pair -s 3 -c 1 -o $OUT/r01_k2_pair_v4_full $CMD > $OUT/ncu_k2_v4_full.log 2>&1
echo "ncu k2 exit $?"; tail -2 $OUT/ncu_k2_v4_full.log
timeout 300 $CMD > $OUT/plain_k2_v4b.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/r01_launches_c2b_v4.csv $CMD > $OUT/ncu_launches_c2b_v4.log 2>&1
echo "ncu launches exit $?"
