#!/bin/bash
# round 2, third GPU pass: same-format 16-bit operands, bf16 multi-query stream kernel; then the 4-CTA-cluster (shared operand) A/B
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python __graft_entry__.py --smoke > $OUT/r2c_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/r2c_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/r2c_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2c_pytest.log
tail -8 $OUT/r2c_pytest.log
for wl in c2b c2bb c4 c5; do
  timeout 600 python bench.py --workload $wl --no-extra --steps 20 --warmup 3 > $OUT/r2c_bench_$wl.json 2> $OUT/r2c_bench_$wl.err; echo "bench $wl exit $?"
done
# experimental: clusters of two pairs sharing an operand by TMA multicast (guarded: a protocol error would hang)
RAGERA_K2_CLUSTER=1 timeout 240 python -m pytest tests/test_gpu_tensor.py -m gpu -x -q > $OUT/r2c_cluster_tests.log 2>&1; rc=$?; echo "cluster tensor tests exit $rc" | tee -a $OUT/r2c_cluster_tests.log
tail -5 $OUT/r2c_cluster_tests.log
if [ $rc -eq 0 ]; then
  for wl in c2b c4 c5; do
    RAGERA_K2_CLUSTER=1 timeout 300 python bench.py --workload $wl --no-extra --steps 20 --warmup 3 > $OUT/r2c_bench_${wl}_cluster.json 2> $OUT/r2c_bench_${wl}_cluster.err; echo "bench $wl cluster exit $?"
  done
fi
