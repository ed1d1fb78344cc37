#!/bin/bash
# One GPU visit: full ncu captures of the three kernel variants the BASELINE configs launch.
#   opaque (C2): 4096^2 kind 0 defaults   alpha (C3): 4096^2 kind 1, filterbank off   uber (C5): 2048^2 kind 0, uber 4, filterbank off
# Usage (under gpurun): bash tools/gpu_prof3.sh <tag> [lib]
set -u
TAG=${1:-x}; LIB=${2:-}
LIBARG=""; [ -n "$LIB" ] && LIBARG="--lib $LIB"
mkdir -p gpurun_out
timeout 300 python tools/prof_target.py --launches 4 $LIBARG > gpurun_out/${TAG}_ms_opaque.txt 2>&1; cat gpurun_out/${TAG}_ms_opaque.txt
timeout 300 python tools/prof_target.py --launches 4 --kind 1 --fb 0 $LIBARG > gpurun_out/${TAG}_ms_alpha.txt 2>&1; cat gpurun_out/${TAG}_ms_alpha.txt
timeout 300 python tools/prof_target.py --launches 3 --size 2048 --uber 4 --fb 0 $LIBARG > gpurun_out/${TAG}_ms_uber.txt 2>&1; cat gpurun_out/${TAG}_ms_uber.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_opaque_prof \
    python tools/prof_target.py --launches 2 $LIBARG > gpurun_out/${TAG}_ncu_opaque.log 2>&1
# alpha texture: launches are (classify, opaque encode, alpha encode) x N -> skip to the alpha kernel of the 2nd launch set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 3 -c 1 -f -o gpurun_out/${TAG}_alpha_prof \
    python tools/prof_target.py --launches 2 --kind 1 --fb 0 $LIBARG > gpurun_out/${TAG}_ncu_alpha.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_uber_prof \
    python tools/prof_target.py --launches 2 --size 2048 --uber 4 --fb 0 $LIBARG > gpurun_out/${TAG}_ncu_uber.log 2>&1
ls -la gpurun_out | grep ${TAG}
