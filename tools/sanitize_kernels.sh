#!/bin/bash
# compute-sanitizer over every hand-written kernel (1 GPU).  Logs -> gpurun_out/sanitizer_*.log;
# copy the summaries to profiles/sanitizers/.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python tools/sanitizer_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|sanitizer target done|Error|hazard" gpurun_out/sanitizer_$tool.log | tail -n 8
done
