#!/bin/bash
# compute-sanitizer over a small end-to-end pass of every kernel:  gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python tests/tools/sanitize_small.py > $OUT/sanitize_plain.log 2>&1; echo "plain exit $?" | tee -a $OUT/sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tests/tools/sanitize_small.py > $OUT/sanitize_$tool.log 2>&1; echo "$tool exit $?" | tee -a $OUT/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small|Error|hazard" $OUT/sanitize_$tool.log | head -8
done
