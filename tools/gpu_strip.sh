#!/bin/bash
# Strip-kernel visit: GPU tests, kernel times of both strip versions (ncu, serialised), bench lines of both.  Usage (under gpurun): bash tools/gpu_strip.sh <tag> [notests]
TAG=${1:-s}
mkdir -p gpurun_out
if [ "$2" != notests ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -n 4 gpurun_out/${TAG}_pytest.log
fi
for v in new old; do
  if [ $v = old ]; then export VKT_BCN_OLD_STRIP=1; else unset VKT_BCN_OLD_STRIP; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resize_ --csv --log-file gpurun_out/${TAG}_strip_${v}.csv python tools/resize_bench.py > gpurun_out/${TAG}_strip_${v}.txt 2>&1
  tail -n 2 gpurun_out/${TAG}_strip_${v}.txt
  python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/${TAG}_strip_${v}.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
for r in rows[1:]: print('$v', r[ki][:48], r[gi], r[vi])
PY
done
unset VKT_BCN_OLD_STRIP
for v in new old; do
  if [ $v = old ]; then export VKT_BCN_OLD_STRIP=1; else unset VKT_BCN_OLD_STRIP; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --strong "" --no-cpu-baseline > gpurun_out/${TAG}_bench_${v}.json 2> gpurun_out/${TAG}_bench_${v}.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_${v}.json').read().strip().split('\n')[-1]); print('$v value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'pageable', d['e2e_pageable']['ms_per_step'], 'devdst', d['e2e_device_destinations']['ms_per_step'], 'dropin', d['e2e_dropin']['ms_per_step'], 'launches', d['gpu_launches'])"
done
unset VKT_BCN_OLD_STRIP
