#!/bin/bash
# Quick GPU visit: chain / drop-in parity, one bench line, the e2e device timeline.  Usage (under gpurun): bash tools/gpu_quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json,sys; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
timeout 300 python tools/trace_e2e.py > gpurun_out/${TAG}_trace.txt 2>&1; tail -45 gpurun_out/${TAG}_trace.txt
