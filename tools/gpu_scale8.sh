#!/bin/bash
# 8-GPU visit: torchrun bench at N = 8 and 4 (weak scaling, one chain per rank), multi-device context tests, and ONE
# 16384^2 chain split over 1 / 8 devices of one context.   Usage (gpurun --gpus 8): bash tools/gpu_scale8.sh <tag>
set -u
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_smi.txt 2>&1
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  echo "bench n$N exit $?"; cut -c1-160 gpurun_out/${TAG}_bench_n$N.json; tail -2 gpurun_out/${TAG}_bench_n$N.err
done
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_multi.log
timeout 900 python tools/multi_ctx_bench.py 16384 > gpurun_out/${TAG}_multictx.txt 2>&1; cat gpurun_out/${TAG}_multictx.txt
