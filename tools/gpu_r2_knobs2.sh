#!/bin/bash
# K2 pipeline depth (RAGERA_K2_STAGES; 6 and 7 only without the selection lists, mode 2) and lockstep on C2b, C4, C5.
OUT=gpurun_out; mkdir -p $OUT
run() {  # name workload env...
  local name=$1 wl=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $wl --no-extra --steps 20 --warmup 5 > $OUT/knob2_$name.json 2> $OUT/knob2_$name.err
  python -c "import json; j=json.load(open('$OUT/knob2_$name.json')); print('$name', 'K2 %.4f ms'%j['kernel_ms_per_step'].get('tensor',0), 'step %.4f'%j['ms_per_step'], 'cert', j.get('certified'))"
}
run c2b_default c2b A=1
for s in 3 4 5; do run c2b_stages$s c2b RAGERA_K2_STAGES=$s; done
for s in 5 6 7; do run c2b_m2_stages$s c2b RAGERA_K2_MODE=2 RAGERA_K2_STAGES=$s; done
for s in 5 7; do run c2b_m2_l0_stages$s c2b RAGERA_K2_MODE=2 RAGERA_K2_STAGES=$s RAGERA_K2_LOCKSTEP=0; done
run c4_default c4 A=1
run c4_lock0 c4 RAGERA_K2_LOCKSTEP=0
run c5_default c5 A=1
run c5_lock0 c5 RAGERA_K2_LOCKSTEP=0
