#!/bin/bash
# Short round-end visit after a change to the opaque encode kernels: all GPU tests, smoke, the default bench (with strong C5 / C3),
# kernel times of the three workloads, full ncu captures of the opaque-default and the uber-4 kernel.  Usage (under gpurun): bash tools/gpu_final3.sh <tag>
set -u
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -n 3 gpurun_out/${TAG}_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -n 1 gpurun_out/${TAG}_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 100 python tools/prof_target.py --launches 4 > gpurun_out/${TAG}_ms_opaque.txt 2>&1; tail -n 1 gpurun_out/${TAG}_ms_opaque.txt
timeout 100 python tools/prof_target.py --launches 3 --size 2048 --uber 4 --fb 0 > gpurun_out/${TAG}_ms_uber.txt 2>&1; tail -n 1 gpurun_out/${TAG}_ms_uber.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_opaque_prof \
    python tools/prof_target.py --launches 2 > gpurun_out/${TAG}_ncu_opaque.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bc7_encode -s 2 -c 1 -f -o gpurun_out/${TAG}_uber_prof \
    python tools/prof_target.py --launches 2 --size 2048 --uber 4 --fb 0 > gpurun_out/${TAG}_ncu_uber.log 2>&1
for k in opaque uber; do ncu -i gpurun_out/${TAG}_${k}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_${k}_raw.csv 2>/dev/null; ncu -i gpurun_out/${TAG}_${k}_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${k}_sass.csv 2>/dev/null; done
ls gpurun_out | grep ${TAG} | wc -l
