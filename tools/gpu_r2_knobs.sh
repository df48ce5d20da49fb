#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
RAGERA_K2_PROF=1 timeout 300 python bench.py --workload c2b --no-extra --steps 20 --warmup 5 > $OUT/knob_c2b_prof.json 2> $OUT/knob_c2b_prof.err; grep -A8 "k2 pair prof" $OUT/knob_c2b_prof.err | head -9
RAGERA_K2_PROF=1 RAGERA_K2_MODE=2 timeout 300 python bench.py --workload c2b --no-extra --steps 20 --warmup 5 > $OUT/knob_c2b_prof_mode2.json 2> $OUT/knob_c2b_prof_mode2.err; grep -A4 "k2 pair prof" $OUT/knob_c2b_prof_mode2.err | head -5
for kv in "RAGERA_K2_PREFETCH=24" "RAGERA_K2_PREFETCH=0" "RAGERA_K2_LOCKSTEP=2" "RAGERA_K2_LOCKSTEP=0"; do
  env $kv timeout 300 python bench.py --workload c2b --no-extra --steps 30 --warmup 5 > $OUT/knob_c2b_$kv.json 2> $OUT/knob_c2b_$kv.err
  python -c "import json,sys; j=json.load(open('$OUT/knob_c2b_$kv.json')); print('$kv', 'K2 %.4f ms'%j['kernel_ms_per_step']['tensor'], 'step %.4f'%j['ms_per_step'])"
done
