#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
RAGERA_K2_PROF=1 timeout 200 python bench.py --workload c2b --no-extra --steps 20 --warmup 5 > $OUT/c2b_prof.json 2> $OUT/c2b_prof.err
echo "exit $?"; grep -A 12 "k2 prof" $OUT/c2b_prof.err | head -60
