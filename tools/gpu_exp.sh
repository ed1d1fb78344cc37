#!/bin/bash
# Time tuning variants (variants/<name>.so) on the three kernel workloads; the sha1 of the blocks must equal the base's.
# Usage (under gpurun): bash tools/gpu_exp.sh <tag> name1 name2 ...
TAG=$1; shift
mkdir -p gpurun_out
for n in "$@"; do
  f=variants/$n.so
  a=$(timeout 200 python tools/prof_target.py --launches 5 --lib $f 2>&1 | tail -1)
  b=$(timeout 200 python tools/prof_target.py --launches 5 --kind 1 --fb 0 --lib $f 2>&1 | tail -1)
  c=$(timeout 200 python tools/prof_target.py --launches 4 --size 2048 --uber 4 --fb 0 --lib $f 2>&1 | tail -1)
  echo "$n | opaque: $a"; echo "$n | alpha : $b"; echo "$n | uber  : $c"
done 2>&1 | sed -E "s/'[0-9.]+', //" | tee gpurun_out/${TAG}_exp.txt
