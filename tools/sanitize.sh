#!/bin/bash
# compute-sanitizer over every kernel (under gpurun): bash tools/sanitize.sh <tag>
T=${1:-x}; mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== $tool" >> gpurun_out/${T}_sanitizer.txt
  VKT_SAN_BIG=$([ $tool = memcheck ] && echo 1) VKT_SAN_NO_THREADS=$([ $tool != memcheck ] && echo 1) timeout ${VKT_SAN_TIMEOUT:-600} compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | tail -3 >> gpurun_out/${T}_sanitizer.txt
done
cat gpurun_out/${T}_sanitizer.txt
