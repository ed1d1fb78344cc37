#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list, one full ncu capture of the level-0 kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
set -u
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [ "${2:-}" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json
timeout 300 python tools/prof_target.py --launches 5 > gpurun_out/${TAG}_kernel_ms.txt 2>&1; cat gpurun_out/${TAG}_kernel_ms.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_prof \
    python tools/prof_target.py --launches 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu exit $?"
ls -la gpurun_out
