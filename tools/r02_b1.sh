#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DINO_B200_GRAPH=0 DINO_B200_PDL=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_b1_launches.csv python tools/b1_launches.py vitl14 1 > gpurun_out/b1.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_b1_launches.csv')) if len(r) > 5]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
seq = []
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    us = v / 1000 if u in ('ns', 'nsecond') else v
    seq.append((r[ki][:60], us))
print(len(seq), 'launches, sum', round(sum(u for _, u in seq), 1), 'us')
# one encoder block in launch order (block 5)
agg = collections.OrderedDict()
for k, u in seq:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += u
for k, (n, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{u:9.1f} us  {n:4d} x {u/n:7.2f}  {k}')
print('--- launches 40..54 in order')
for k, u in seq[40:54]:
    print(f'{u:8.2f}  {k}')
PY
