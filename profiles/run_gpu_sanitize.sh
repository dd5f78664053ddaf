#!/bin/bash
# compute-sanitizer over small cases of every kernel family (tools/sanitize_cases.py): memcheck, racecheck (shared-memory
# hazards of the barrier-free mbarrier rings and of the ghost warps) and synccheck.  Logs -> gpurun_out/<tag>/, the
# summaries are copied to profiles/.   usage: bash profiles/run_gpu_sanitize.sh <tag>
TAG=${1:-sanitize}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_cases.py > $OUT/sanitizer_$tool.log 2>&1
  echo "$tool exit code $?" >> $OUT/sanitizer_$tool.log
  tail -4 $OUT/sanitizer_$tool.log
done
