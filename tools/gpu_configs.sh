#!/bin/bash
# One GPU visit over the BASELINE configs: parity tests, default bench, c3/c5 bench lines, ncu captures of the alpha
# kernel and of the uber-4 kernel.   Usage (under gpurun): bash tools/gpu_configs.sh <tag> [skip-tests]
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
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
echo "bench c3 exit $?"; cat gpurun_out/${TAG}_bench_c3.json; tail -3 gpurun_out/${TAG}_bench_c3.err
timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
echo "bench c5 exit $?"; cat gpurun_out/${TAG}_bench_c5.json; tail -3 gpurun_out/${TAG}_bench_c5.err
timeout 300 python tools/prof_target.py --size 4096 --kind 1 --fb 0 --launches 4 > gpurun_out/${TAG}_kernel_ms_alpha.txt 2>&1; cat gpurun_out/${TAG}_kernel_ms_alpha.txt
timeout 300 python tools/prof_target.py --size 4096 --kind 0 --uber 4 --fb 0 --launches 3 > gpurun_out/${TAG}_kernel_ms_uber4.txt 2>&1; cat gpurun_out/${TAG}_kernel_ms_uber4.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 3 -c 1 -f -o gpurun_out/${TAG}_alpha_prof \
    python tools/prof_target.py --size 4096 --kind 1 --fb 0 --launches 2 > gpurun_out/${TAG}_ncu_alpha.log 2>&1
echo "ncu alpha exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_uber4_prof \
    python tools/prof_target.py --size 2048 --kind 0 --uber 4 --fb 0 --launches 2 > gpurun_out/${TAG}_ncu_uber4.log 2>&1
echo "ncu uber4 exit $?"
ls -la gpurun_out
