#!/bin/bash
# compute-sanitizer over small batches of every hot kernel (memcheck, racecheck on shared memory, synccheck)
mkdir -p gpurun_out
: > gpurun_out/sanitizer.log
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> gpurun_out/sanitizer.log
  timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $tool python scripts/sanitize_small.py 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize_small|ERROR SUMMARY|Error|error|hazard" | head -20 >> gpurun_out/sanitizer.log
done
cat gpurun_out/sanitizer.log
