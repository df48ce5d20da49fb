#!/bin/bash
# K2 bring-up: run the tensor-path tests under a hard timeout so a hung kernel cannot hold the box.
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_tensor.py -x -q "$@" > $OUT/k2_pytest.log 2>&1; echo "k2 pytest exit $?" | tee -a $OUT/k2_pytest.log
tail -40 $OUT/k2_pytest.log
