#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python bench.py --workload c2b --no-extra --steps 50 --warmup 5 > $OUT/c2b_bench.json 2> $OUT/c2b_bench.err; echo "c2b exit $?"
cat $OUT/c2b_bench.json | head -c 3000; tail -5 $OUT/c2b_bench.err
BENCH_SHORT="python bench.py --workload c2 --steps 5 --warmup 3 --no-extra"
timeout 300 $BENCH_SHORT > $OUT/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 5 -c 2 -o $OUT/r01_k1_full $BENCH_SHORT > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 $OUT/ncu_full.log
