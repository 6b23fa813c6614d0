#!/bin/bash
# compute-sanitizer passes over the tiny end-to-end path (vision -> prefill -> graph-less + graph decode -> generate)
set -u
mkdir -p gpurun_out
export WATCHDOG=600
for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck}; do
  timeout ${SANITIZE_TIMEOUT:-400} compute-sanitizer --tool $tool --launch-timeout 0 --error-exitcode 9 python tools/decode_debug.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|DONE" gpurun_out/sanitize_$tool.log | tail -5 | tee -a gpurun_out/sanitize_summary.log
done
