#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): multi-device context tests, the torchrun bench at N and at 1, the reference arm
# under torchrun.   Usage: bash tools/gpu_multi.sh <tag> <N>
set -u
TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench n1 exit $?"; cat gpurun_out/${TAG}_bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench n$N exit $?"; cat gpurun_out/${TAG}_bench_n${N}.json; tail -5 gpurun_out/${TAG}_bench_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n${N}.json 2> gpurun_out/${TAG}_bench_ref_n${N}.err
echo "ref n$N exit $?"; cat gpurun_out/${TAG}_bench_ref_n${N}.json; tail -3 gpurun_out/${TAG}_bench_ref_n${N}.err
timeout 600 python tools/multi_ctx_bench.py > gpurun_out/${TAG}_multictx.txt 2>&1; cat gpurun_out/${TAG}_multictx.txt
