"""Hottest source lines of a kernel from an ncu report captured with --import-source on (development aid).

    python tools/ncu_lines.py report.ncu-rep <kernel regex> [top N] [launch index]
Prints per CUDA source line: warp instructions executed (and share), stall samples, average active threads.
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
# the export holds one block per (launch, file); a launch starts with a "Kernel Name"-less "File Path" header sequence
blocks, cur, fname = [], None, None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1]
        continue
    if row[0] == "Function Name":
        func = row[1]
        continue
    if row[0] == "Line No":
        hdr = row
        cur = dict(file=fname, func=func, hdr=hdr, rows=[])
        blocks.append(cur)
        continue
    if cur is not None:
        cur["rows"].append(row)
# group blocks into launches: a new launch begins when the first file repeats
launches, seen = [], set()
for b in blocks:
    key = (b["file"], b["func"])
    if not launches or key in seen:
        launches.append([])
        seen = set()
    seen.add(key)
    launches[-1].append(b)
print("launches in report for this kernel:", len(launches))
L = launches[which]
lines = []
for b in L:
    h = b["hdr"]
    ia, ii, isamp, ithr = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
    for r in b["rows"]:
        if len(r) <= ithr or r[ia] != "-":
            continue
        try:
            n, s, t = int(r[ii]), int(r[isamp]), int(r[ithr])
        except ValueError:
            continue
        if n:
            lines.append((n, s, t, b["file"].split("/")[-1], r[0], r[1].strip()))
tot = sum(x[0] for x in lines)
tots = sum(x[1] for x in lines)
print("total warp instr %.1f M, samples %d" % (tot / 1e6, tots))
for n, s, t, f, ln, src in sorted(lines, reverse=True)[:top]:
    print("%6.2f%% instr %5.1f%% samp  thr %4.1f  %s:%s  %s" % (100.0 * n / tot, 100.0 * s / max(tots, 1), t / n, f, ln, src[:110]))
