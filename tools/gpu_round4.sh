#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_batcher.py tests/test_loader.py -m gpu -x -q > $OUT/r4_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r4_pytest.log
CMD="python bench.py --workload c2b --no-extra --steps 4 --warmup 3"
timeout 300 $CMD > $OUT/plain_k2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_pair -s 3 -c 1 -o $OUT/r01_k2_pair_v2_full $CMD > $OUT/ncu_k2_full.log 2>&1
echo "ncu k2 exit $?"; tail -2 $OUT/ncu_k2_full.log
timeout 300 $CMD > $OUT/plain_k2b.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/r01_launches_c2b.csv $CMD > $OUT/ncu_launches_c2b.log 2>&1
echo "ncu launches exit $?"
