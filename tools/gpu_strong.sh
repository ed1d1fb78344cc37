#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): torchrun bench at N (headline replicas + strong C5 / C3 chains), optional tests.
# Usage: bash tools/gpu_strong.sh <tag> <N> [tests]
set -u
TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
if [ "${3:-}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench n$N exit $?"; tail -5 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
for k,v in d.get("strong",{}).items():
    print(k, {a:b for a,b in v.items() if a not in ("workload","partitioning","timing")})
PY
