#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_napi.py tests/test_gpu_batcher.py -m gpu -x -q > $OUT/r2h_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2h_pytest.log
tail -4 $OUT/r2h_pytest.log
for wl in c1 c2; do
  RAGERA_SMALL_PROF=1 timeout 300 python bench.py --workload $wl --no-extra --steps 500 --warmup 20 > $OUT/r2h_bench_${wl}_prof.json 2> $OUT/r2h_bench_${wl}_prof.err; echo "$wl prof exit $?"
  grep "prof\]" $OUT/r2h_bench_${wl}_prof.err
  timeout 300 python bench.py --workload $wl --no-extra --steps 1000 --warmup 20 > $OUT/r2h_bench_${wl}.json 2> $OUT/r2h_bench_${wl}.err; echo "$wl exit $?"
done
