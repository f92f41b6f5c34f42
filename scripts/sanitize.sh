#!/bin/bash
# compute-sanitizer over small invocations of every kernel path (race detection / sanitizers), to be run on the
# GPU box:  bash scripts/sanitize.sh   -> gpurun_out/sanitize_{memcheck,racecheck,synccheck}.log
# (The round-1 pool refuses compute-sanitizer - "closed on this pool" - so only the plain all-paths run of the driver was
# exercised there; the script is kept for boxes where the tool is available.)
mkdir -p gpurun_out
python scripts/sanitize_driver.py 24 > gpurun_out/sanitize_plain.log 2>&1 || { echo "driver failed without the sanitizer"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_driver.py 6 > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool: exit $? | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
