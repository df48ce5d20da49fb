#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q -s > $OUT/r2f_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2f_pytest.log
tail -5 $OUT/r2f_pytest.log
RAGERA_SMALL_PROF=1 timeout 300 python bench.py --workload c1 --no-extra --steps 500 --warmup 20 > $OUT/r2f_bench_c1_prof.json 2> $OUT/r2f_bench_c1_prof.err; echo "c1 prof exit $?"
grep "prof\]" $OUT/r2f_bench_c1_prof.err
RAGERA_SMALL_PROF=1 timeout 300 python bench.py --workload c2 --no-extra --steps 50 --warmup 5 > $OUT/r2f_bench_c2_prof.json 2> $OUT/r2f_bench_c2_prof.err; echo "c2 prof exit $?"
grep "prof\]" $OUT/r2f_bench_c2_prof.err
bash tools/gpu_ncu.sh c3 k1_stream_f32
bash tools/gpu_ncu.sh c2b k2_pair_kernel
