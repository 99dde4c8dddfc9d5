"""SASS instructions with the most stall samples for a kernel of an ncu report (development aid).
    python tools/ncu_sass_hot.py report.ncu-rep <kernel regex> [min share %] [launch index]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
        continue
    if r and r[0] == "Address":
        cur["hdr"] = r
        continue
    if cur is not None and r:
        cur["rows"].append(r)
b = blocks[which]
h = b["hdr"]
isamp, iex, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[isamp]) for r in b["rows"])
print(b["name"][:90], "launches:", len(blocks), "samples:", tot)
for i, r in enumerate(b["rows"]):
    n = int(r[isamp])
    if n > tot * minshare / 100:
        top = sorted(((int(r[j]), h[j]) for j in stalls), reverse=True)[0]
        print("%5d %s %-72s %5.1f%% (%s %d) exec %s" % (i, r[0][-5:], r[isrc].strip()[:72], 100.0 * n / tot, top[1], top[0], r[iex]))
