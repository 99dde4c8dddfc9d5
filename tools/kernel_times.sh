#!/bin/bash
# per-kernel time and warp instructions of one 8K encode + decode (development aid): tools/kernel_times.sh [kernel regex]
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"${1:-.}" --csv python tests/gpu_dec_once.py 7680 4320 1 2>/dev/null | python -c "
import csv, sys, collections
rows = [r for r in csv.reader(sys.stdin) if len(r) > 10]
if not rows: sys.exit('no rows')
h = rows[0]; ik, im, iv = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ik].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    a = agg.setdefault(n, [0, 0.0, 0.0])
    v = float(r[iv].replace(',', ''))
    if r[im] == 'gpu__time_duration.sum': a[0] += 1; a[1] += v / 1e3
    else: a[2] += v / 1e6
for n, (c, t, i) in agg.items(): print('%-34s x%-2d %8.1f us %8.1f M warp instr' % (n[:34], c, t, i))
"
