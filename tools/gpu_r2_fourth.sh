#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/r2d_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2d_pytest.log
tail -8 $OUT/r2d_pytest.log
# experimental: clusters of two pairs sharing an operand by TMA multicast (guarded: a protocol error would hang)
RAGERA_K2_CLUSTER=1 timeout 300 python -m pytest tests/test_gpu_tensor.py tests/test_certification.py -m gpu -x -q > $OUT/r2d_cluster_tests.log 2>&1; rc=$?; echo "cluster tensor tests exit $rc" | tee -a $OUT/r2d_cluster_tests.log
tail -5 $OUT/r2d_cluster_tests.log
if [ $rc -eq 0 ]; then
  for wl in c2b c4 c5; do
    RAGERA_K2_CLUSTER=1 timeout 300 python bench.py --workload $wl --no-extra --steps 20 --warmup 3 > $OUT/r2d_bench_${wl}_cluster.json 2> $OUT/r2d_bench_${wl}_cluster.err; echo "bench $wl cluster exit $?"
  done
fi
