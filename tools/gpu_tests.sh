#!/bin/bash
# A subset of the GPU suite, optionally under environment knobs:
#   gpurun -- 'RAGERA_K2_CLUSTER=2 bash tools/gpu_tests.sh cluster2 tests/test_gpu_tensor.py tests/test_certification.py'
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest "$@" -m gpu -x -q -s > $OUT/pytest_$TAG.log 2>&1; echo "pytest ($TAG) exit $?" | tee -a $OUT/pytest_$TAG.log
tail -6 $OUT/pytest_$TAG.log
