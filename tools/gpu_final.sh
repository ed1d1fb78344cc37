#!/bin/bash
# Round-end visit on one GPU: all GPU tests, smoke, the default bench (with strong C5 / C3), the reference arm, the ncu launch
# list of the bench command and full ncu captures of the three kernel variants the BASELINE configs launch.
# Usage (under gpurun): bash tools/gpu_final.sh <tag>
set -u
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -1 gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref exit $?"; cat gpurun_out/${TAG}_bench_ref.json | cut -c1-400
for wl in c1 c3 c5; do timeout 600 python bench.py --workload $wl --strong "" --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err; echo "bench $wl exit $?"; done
timeout 600 python bench.py --workload c4 --strong "" --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; echo "bench c4 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong "" > gpurun_out/${TAG}_ncu_bench.log 2>&1
bash tools/gpu_prof3.sh ${TAG} > gpurun_out/${TAG}_prof3.log 2>&1; head -3 gpurun_out/${TAG}_prof3.log
timeout 200 python tools/dropin_e2e.py > gpurun_out/${TAG}_dropin.txt 2>&1; tail -1 gpurun_out/${TAG}_dropin.txt
timeout 200 python tools/pageable_e2e.py > gpurun_out/${TAG}_pageable.txt 2>&1; tail -4 gpurun_out/${TAG}_pageable.txt
timeout 200 python tools/e2e_sweep.py > gpurun_out/${TAG}_sweep.txt 2>&1; tail -5 gpurun_out/${TAG}_sweep.txt
ls gpurun_out | grep ${TAG} | wc -l
