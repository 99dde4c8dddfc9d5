"""List the loops (backward branches) of a kernel's SASS with their instruction mix -- a development aid.

    cuobjdump -sass file.o | python tools/sass_loops.py <function substring> [min_instr]
"""
import collections
import re
import sys

want = sys.argv[1]
min_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ins = []
on = False
for line in sys.stdin:
    if "Function :" in line:
        on = want in line
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)*(0x[0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt < a and tgt in addr_idx:
        j = addr_idx[tgt]
        n = i - j + 1
        if n < min_n:
            continue
        hist = collections.Counter()
        for _, tt in ins[j:i + 1]:
            tt = re.sub(r"^@!?U?P\d+\s+", "", tt)
            hist[tt.split()[0].split(".")[0]] += 1
        print("loop 0x%x..0x%x: %d instr  %s" % (tgt, a, n, dict(hist.most_common())))
