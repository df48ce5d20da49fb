#!/bin/bash
# round 2, first GPU pass: smoke, the GPU test suite, a short default bench
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/r2_gpu.txt 2>&1
nproc >> $OUT/r2_gpu.txt; free -g | head -2 >> $OUT/r2_gpu.txt
timeout 300 python __graft_entry__.py --smoke > $OUT/r2_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=15 > $OUT/r2_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2_pytest.log
tail -30 $OUT/r2_pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 > $OUT/r2_bench_default.json 2> $OUT/r2_bench_default.err; echo "bench exit $?"
tail -c 1500 $OUT/r2_bench_default.err
