#!/bin/bash
# kernel list of one latency-mode solve under ncu (cold-cache, serialised durations): tools/gpu_lat_ncu.sh rows cols cell bins
mkdir -p gpurun_out
NID_SOLVE_ONLY=1 NID_OPTS=${NID_OPTS:-lm_graph=0} timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/lat_launches.csv \
  python tools/time_single.py "$@" > gpurun_out/lat_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/lat_launches.csv')) if len(r) > 10]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); ig = hdr.index('Grid Size'); ib = hdr.index('Block Size')
seq = [(r[ik].split('(')[0], r[ig], r[ib], float(r[iv].replace(',', '')) / 1000.0) for r in rows[1:]]
for k, g, b, v in seq[-12:]: print('%-34s %-16s %-14s %7.1f us' % (k, g, b, v))
PY
