#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
RAGERA_K2_CLUSTER=2 timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_certification.py -m gpu -x -q > $OUT/r2g_cluster2_tests.log 2>&1; rc=$?; echo "cluster=2 tensor tests exit $rc" | tee -a $OUT/r2g_cluster2_tests.log
tail -5 $OUT/r2g_cluster2_tests.log
if [ $rc -eq 0 ]; then
  for c in 2 0; do
    RAGERA_K2_CLUSTER=$c bash tools/gpu_bench.sh cluster$c c2b c4 c5
  done
fi
