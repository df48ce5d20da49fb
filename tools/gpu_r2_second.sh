#!/bin/bash
# round 2, second GPU pass: fp16 operands — tensor/certification tests first, then the suite, then K2 mode A/B on c2b
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_certification.py -m gpu -x -q > $OUT/r2b_tensor.log 2>&1; echo "tensor tests exit $?" | tee -a $OUT/r2b_tensor.log
tail -15 $OUT/r2b_tensor.log
timeout 300 python __graft_entry__.py --smoke > $OUT/r2b_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/r2b_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/r2b_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2b_pytest.log
tail -8 $OUT/r2b_pytest.log
for wl in c2b c2bb c4 c5; do
  timeout 600 python bench.py --workload $wl --no-extra --steps 20 --warmup 3 > $OUT/r2b_bench_$wl.json 2> $OUT/r2b_bench_$wl.err; echo "bench $wl exit $?"
done
for mode in 1 2; do
  RAGERA_K2_MODE=$mode timeout 300 python bench.py --workload c2b --no-extra --steps 20 --warmup 3 > $OUT/r2b_bench_c2b_mode$mode.json 2> $OUT/r2b_bench_c2b_mode$mode.err; echo "bench c2b mode $mode exit $?"
done
RAGERA_K2_QFMT=bf16 timeout 300 python bench.py --workload c5 --no-extra --steps 10 --warmup 3 > $OUT/r2b_bench_c5_qbf16.json 2> $OUT/r2b_bench_c5_qbf16.err; echo "bench c5 qbf16 exit $?"
