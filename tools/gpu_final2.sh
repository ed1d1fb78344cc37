#!/bin/bash
# Round-end visit after a change outside the encode kernels (their ncu captures stay valid): all GPU tests, smoke, the default
# bench (with strong C5 / C3), the reference arm, the other configs, the ncu launch list of the bench command, full ncu captures
# of the two resize strip kernels, drop-in / pageable / size sweeps.   Usage (under gpurun): bash tools/gpu_final2.sh <tag>
set -u
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -n 3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -n 1 gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
for wl in c1 c3 c4 c5; do timeout 600 python bench.py --workload $wl --strong "" --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err; echo "bench $wl exit $?"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong "" > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resize_strip -s 2 -c 1 -f -o gpurun_out/${TAG}_strip11_prof python tools/resize_bench.py > gpurun_out/${TAG}_ncu_strip11.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resize_strip -s 5 -c 1 -f -o gpurun_out/${TAG}_strip21_prof python tools/resize_bench.py > gpurun_out/${TAG}_ncu_strip21.log 2>&1
for k in strip11 strip21; do ncu -i gpurun_out/${TAG}_${k}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_${k}_raw.csv 2>/dev/null; done
timeout 200 python tools/dropin_e2e.py > gpurun_out/${TAG}_dropin.txt 2>&1; tail -n 1 gpurun_out/${TAG}_dropin.txt
timeout 200 python tools/pageable_e2e.py > gpurun_out/${TAG}_pageable.txt 2>&1; tail -n 4 gpurun_out/${TAG}_pageable.txt
timeout 200 python tools/e2e_sweep.py > gpurun_out/${TAG}_sweep.txt 2>&1; tail -n 5 gpurun_out/${TAG}_sweep.txt
timeout 200 python tools/trace_e2e.py > gpurun_out/${TAG}_trace.txt 2>&1
ls gpurun_out | grep ${TAG} | wc -l
for st in 8 32; do VKT_BCN_RESIZE_STRIP=$st timeout 300 python bench.py --steps 20 --warmup 5 --strong "" --no-cpu-baseline > gpurun_out/${TAG}_bench_strip$st.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_strip$st.json').read().strip().split('\n')[-1]); print('strip $st e2e', d['e2e']['ms_per_step'], 'pageable', d['e2e_pageable']['ms_per_step'])"; done
