#!/bin/bash
# One bench line per workload (environment knobs pass through):  gpurun -- 'RAGERA_GRAPH=0 bash tools/gpu_bench.sh tag c1 c2b'
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for wl in "$@"; do
  timeout 900 python bench.py --workload $wl --no-extra ${BENCH_ARGS:---steps 30 --warmup 5} > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err; echo "bench $wl ($TAG) exit $?"
done
