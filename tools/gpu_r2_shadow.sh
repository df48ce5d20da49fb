#!/bin/bash
# round 2, shadow stream path: smoke, its tests + the sharded-on-one-GPU test, then the bench lines c3s / c1s next to c3 / c1
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python __graft_entry__.py --smoke > $OUT/ss_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/ss_smoke.log; tail -2 $OUT/ss_smoke.log
timeout 900 python -m pytest tests/test_gpu_shadow_stream.py tests/test_gpu_sharded.py -x -q -s > $OUT/ss_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/ss_pytest.log
tail -15 $OUT/ss_pytest.log
for w in c3s c1s c1; do
  timeout 600 python bench.py --workload $w --no-extra > $OUT/ss_bench_$w.json 2> $OUT/ss_bench_$w.err; echo "bench $w exit $?"
done
