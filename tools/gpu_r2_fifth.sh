#!/bin/bash
# graph replay for the small-batch path, DADD latency, C1 A/B
OUT=gpurun_out; mkdir -p $OUT
./tools/micro/dadd_latency > $OUT/r2_dadd_latency.txt 2>&1; cat $OUT/r2_dadd_latency.txt
timeout 300 python __graft_entry__.py --smoke > $OUT/r2e_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/r2e_smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q -s > $OUT/r2e_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2e_pytest.log
tail -6 $OUT/r2e_pytest.log
for g in 1 0; do
  RAGERA_GRAPH=$g timeout 300 python bench.py --workload c1 --no-extra --steps 2000 --warmup 50 > $OUT/r2e_bench_c1_graph$g.json 2> $OUT/r2e_bench_c1_graph$g.err; echo "bench c1 graph=$g exit $?"
  RAGERA_GRAPH=$g timeout 300 python bench.py --workload c2 --no-extra --steps 200 --warmup 10 > $OUT/r2e_bench_c2_graph$g.json 2> $OUT/r2e_bench_c2_graph$g.err; echo "bench c2 graph=$g exit $?"
done
