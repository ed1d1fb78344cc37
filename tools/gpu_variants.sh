#!/bin/bash
# time every variants/*.so (and the product .so) on the level-0 kernel
mkdir -p gpurun_out
for kind in 0 1; do
echo "== kind $kind product"; python tools/prof_target.py --launches 4 --kind $kind 2>&1 | tail -1
for f in variants/*.so; do echo "== kind $kind $f"; python tools/prof_target.py --launches 4 --kind $kind --lib $f 2>&1 | tail -1; done
done 2>&1 | tee gpurun_out/variants.txt
