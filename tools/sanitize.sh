#!/bin/bash
# compute-sanitizer over every kernel of the library (tools/sanitize_cases.py); logs under gpurun_out/ (copied to profiles/ when clean).
# usage (on the GPU box): bash tools/sanitize.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool"
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/${tag}_sanitizer_$tool.txt 2>&1
  echo "rc=$?" >> gpurun_out/${tag}_sanitizer_$tool.txt
  tail -4 gpurun_out/${tag}_sanitizer_$tool.txt
done
