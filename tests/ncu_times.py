"""Sum gpu__time_duration per kernel from `ncu --metrics gpu__time_duration.sum` text output on stdin (development aid)."""
import re
import sys
import collections

agg = collections.OrderedDict()
name = None
for line in sys.stdin:
    m = re.match(r"^\s+(?:void )?(?:<unnamed>::)?([A-Za-z_0-9]+)(<[^>]*>)?\(.*\) \((\d+), (\d+), (\d+)\)x\((\d+)", line)
    if m:
        name = m.group(1) + (m.group(2) or "")
        continue
    m = re.match(r"^\s+gpu__time_duration.sum\s+(\w+)\s+([\d.,]+)", line)
    if m and name:
        v = float(m.group(2).replace(",", ""))
        u = m.group(1)
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        name = None
tot = sum(t for _, t in agg.values())
for k, (n, t) in agg.items():
    print("%-34s n=%4d %10.1f us %5.1f%%" % (k, n, t, 100 * t / tot if tot else 0))
print("%-34s        %10.1f us" % ("total", tot))
