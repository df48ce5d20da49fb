#!/bin/bash
# lockstep poll moved to warp 2: parity, then K2 time on C2b / C5 with lockstep 1 and 0
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q -m gpu 2>&1 | tail -3
run() {
  local name=$1 wl=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $wl --no-extra --steps 20 --warmup 5 > $OUT/knob3_$name.json 2> $OUT/knob3_$name.err
  python -c "import json; j=json.load(open('$OUT/knob3_$name.json')); print('$name', 'K2 %.4f ms'%j['kernel_ms_per_step'].get('tensor',0), 'step %.4f'%j['ms_per_step'], 'cert', j['certified']['timed_steps'])"
}
run c2b_default c2b A=1
run c2b_lock0 c2b RAGERA_K2_LOCKSTEP=0
run c2b_default_again c2b A=1
run c2b_lock2 c2b RAGERA_K2_LOCKSTEP=2
run c5_default c5 A=1
run c5_lock2 c5 RAGERA_K2_LOCKSTEP=2
RAGERA_K2_PROF=1 timeout 300 python bench.py --workload c2b --no-extra --steps 20 --warmup 5 > $OUT/knob3_c2b_prof.json 2> $OUT/knob3_c2b_prof.err; grep -A8 "k2 pair prof" $OUT/knob3_c2b_prof.err | head -9
