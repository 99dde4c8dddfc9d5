"""Turn ncu exports into the tables of profiles/rNN_summary.md (run here, no GPU needed).

    python profiles/summarize.py launches gpurun_out/r01_launches.csv [frames_total]
    ncu -i gpurun_out/r01_top.ncu-rep --page raw --csv > /tmp/top_raw.csv && python profiles/summarize.py kernels /tmp/top_raw.csv
    python profiles/summarize.py traffic /tmp/top_raw.csv "<source note>" > profiles/traffic.json   # DRAM bytes per stage
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z_0-9]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")).replace("(int)", "") if m else name[:40]


def launches(path, frames):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) != len(hdr) or r == hdr:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
        a = agg.setdefault(short(d["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print("| kernel | launches/frame | us/frame | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %.1f | %.1f | %.1f %% |" % (k, n / frames, t / frames, 100 * t / tot))
    print("| total | %.1f | %.1f | |" % (sum(n for n, _ in agg.values()) / frames, tot / frames))


def kernels(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    want = [("time us", "gpu__time_duration.sum", 1e3), ("DRAM read MB", "dram__bytes_read.sum", 1.0),
            ("DRAM write MB", "dram__bytes_write.sum", 1.0),
            ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
            ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
            ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0),
            ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0),
            ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
            ("warp instr M", "smsp__inst_executed.sum", 1e-6),
            ("threads/instr", "smsp__thread_inst_executed_per_inst_executed.ratio", 1.0),
            ("regs", "launch__registers_per_thread", 1.0),
            ("occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0)]
    unit_row = rows[1]
    print("| kernel (grid) | " + " | ".join(w[0] for w in want) + " |\n|---|" + "---|" * len(want))
    ik, ig = hdr.index("Kernel Name"), hdr.index("Grid Size") if "Grid Size" in hdr else -1
    for r in rows[2:]:
        cells = []
        for label, key, scale in want:
            if key not in hdr:
                cells.append("-")
                continue
            i = hdr.index(key)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                cells.append("-")
                continue
            u = unit_row[i]
            if key == "gpu__time_duration.sum":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
                scale = 1.0
            if key.startswith("dram__bytes"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            cells.append("%.1f" % (v * scale))
        print("| %s %s | %s |" % (short(r[ik]), r[ig] if ig >= 0 else "", " | ".join(cells)))


def stage_of(name):
    n = short(name)
    for prefix, stage in (("lift_fwd", "lift_fwd"), ("lift_tail_fwd", "lift_fwd"), ("lift_inv", "lift_inv"), ("lift_tail_inv", "lift_inv"),
                          ("linearize", "linearize"), ("reconstruct", "reconstruct"), ("enc_", "enc_coder"), ("dec_", "dec_coder")):
        if n.startswith(prefix):
            return stage
    return None


def traffic(path, note):
    """dram__bytes_read.sum + dram__bytes_write.sum of the launches of one frame (raw page of ONE encode + decode), per stage"""
    import json
    rows = list(csv.reader(open(path)))
    hdr, unit_row = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    out = collections.OrderedDict()
    for r in rows[2:]:
        st = stage_of(r[ik])
        if not st:
            continue
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(key)
            v = float(r[i].replace(",", ""))
            tot += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit_row[i], 1.0)
        out[st] = out.get(st, 0.0) + tot
    d = collections.OrderedDict((k, int(v)) for k, v in out.items())
    d["source"] = note
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        kernels(sys.argv[2])
