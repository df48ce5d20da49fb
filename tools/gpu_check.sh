#!/bin/bash
# What the driver runs at round end, in one call:  gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python __graft_entry__.py --smoke > $OUT/check_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/check_smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q -s > $OUT/check_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/check_pytest.log
tail -6 $OUT/check_pytest.log
timeout 1200 python bench.py > $OUT/check_bench.json 2> $OUT/check_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/check_bench_reference.json 2> $OUT/check_bench_reference.err; echo "reference arm exit $?"
