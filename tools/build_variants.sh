#!/bin/bash
# Build tuning variants of the CUDA library into gpurun_variants/ (NOT the product .so): "threads:ctas" pairs.
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for v in "$@"; do
  t=${v%%:*}; c=${v##*:}
  VKT_NVCC_EXTRA="-DVKT_BC7_THREADS=$t -DVKT_BC7_CTAS=$c" VKT_CUDA_SO_OUT=variants/lib_t${t}_c${c}.so \
    python -c "from vierkant_b200 import build; build.build_cuda(force=True, verbose=True)" 2>&1 | grep -A1 "ILb1ELi1" | grep -E "Used|spill" | sed "s/^/t$t c$c: /" &
done
wait
