#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
RAGERA_K2_PROF=1 timeout 400 python bench.py --workload c5 --no-extra --steps 12 --warmup 3 > $OUT/c5_prof.json 2> $OUT/c5_prof.err
grep -A 7 "k2 pair prof" $OUT/c5_prof.err | head -8 | cut -c1-230
CMD="python bench.py --workload c2b --no-extra --steps 4 --warmup 3"
timeout 600 ncu --set full --clock-control none -k regex:k4_rescore -s 3 -c 1 -o $OUT/r01_k4_rescore_full $CMD > $OUT/ncu_k4_full.log 2>&1; echo "ncu k4 exit $?"
