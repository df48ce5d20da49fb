#!/bin/bash
# One gpurun call: GPU test suite, smoke, bench, then the ncu launch list and one full capture
# of the dominant kernel. Everything lands in gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/${TAG}_gpu.csv 2>&1
nproc > $OUT/${TAG}_nproc.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/${TAG}_nproc.txt; free -g >> $OUT/${TAG}_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -c 1500 $OUT/${TAG}_bench.json
BENCH_SHORT="python bench.py --workload c2 --steps 5 --warmup 3 --no-extra"
timeout 300 $BENCH_SHORT > $OUT/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $BENCH_SHORT > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 300 $BENCH_SHORT > $OUT/${TAG}_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 5 -c 2 -o $OUT/${TAG}_k1_full $BENCH_SHORT > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
