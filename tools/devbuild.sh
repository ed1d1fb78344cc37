#!/bin/bash
# Tuning build: the perceptual 28-bit-key kernels only (defaults + uber, opaque + alpha), into variants/<name>.so; prints registers/spills.
# Usage: tools/devbuild.sh <name> [extra nvcc flags]
cd "$(dirname "$0")/.."; mkdir -p variants
N=$1; shift
VKT_NVCC_EXTRA="-DVKT_BC7_DEV_DEFAULT_VARIANTS_ONLY $*" VKT_CUDA_SO_OUT=variants/$N.so \
  python -c "from vierkant_b200 import build; build.build_cuda(force=True, verbose=True)" 2>&1 | grep -A2 "bc7_encode_kernelILb1ELi1" | grep -E "Function properties|Used|spill" | paste - - - | sed 's/ptxas info *: //g' | awk '{print}' | cut -c1-260
