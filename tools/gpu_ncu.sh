#!/bin/bash
# ncu evidence for one workload (B200_PROFILING.md recipe): the launch list of a short bench run, then ONE --set full capture
# of the dominant kernel.   gpurun --timeout 1500 -- 'bash tools/gpu_ncu.sh c2b k2_pair_kernel'
WL=${1:-c3}; KERNEL=${2:-k1_stream_f32}; OUT=gpurun_out; mkdir -p $OUT
ARGS="--workload $WL --no-extra --steps 2 --warmup 3"
timeout 600 python bench.py $ARGS > $OUT/ncu_${WL}_plain.json 2> $OUT/ncu_${WL}_plain.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$WL.csv python bench.py $ARGS > $OUT/ncu_${WL}_list.log 2>&1; echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 6 -c 1 -o $OUT/${WL}_${KERNEL}_full -f python bench.py $ARGS > $OUT/ncu_${WL}_full.log 2>&1; echo "full capture exit $?"
ncu -i $OUT/${WL}_${KERNEL}_full.ncu-rep --page raw --csv > $OUT/${WL}_${KERNEL}_full.raw.csv 2>/dev/null
