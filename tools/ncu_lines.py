#!/usr/bin/env python
"""Per-source-line stall samples / instruction counts from an .ncu-rep (needs -lineinfo and --import-source on).
usage: tools/ncu_lines.py report.ncu-rep [top_n] [kernel-id]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
if len(sys.argv) > 3:
    cmd += ["--kernel-id", sys.argv[3]]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, data = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    iAll = hdr.index("Warp Stall Sampling (All Samples)"); iInst = hdr.index("Instructions Executed")
    def col(name):
        try: return int(r[hdr.index(name)] or 0)
        except Exception: return 0
    try:
        data.append((int(r[iAll] or 0), int(r[iInst] or 0), fname, int(r[0]), r[1].strip(),
                     col("stall_long_sb"), col("stall_short_sb"), col("stall_wait"), col("stall_math"), col("stall_lg"), col("stall_mio"), col("stall_barrier"), col("stall_membar")))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1; toti = sum(d[1] for d in data) or 1
print(f"total samples {tot}  total warp-instructions {toti}")
print(" samp%  inst%  file:line  [long short wait math lg mio bar membar]  source")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100*d[0]/tot:5.1f}  {100*d[1]/toti:5.1f}  {d[2]}:{d[3]:<5} [{d[5]} {d[6]} {d[7]} {d[8]} {d[9]} {d[10]} {d[11]} {d[12]}]  {d[4][:110]}")
