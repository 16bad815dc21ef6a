"""Hot straight-line runs of a kernel's SASS (per-instruction execution counts of an ncu source page).
usage: tools/ncu_sass_runs.py report.ncu-rep kernel-regex [min_share_pct]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name",
                      "regex:" + kre, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
for k, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr, start = r, k + 1
        break
iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [(c, hdr.index(c)) for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
ins = []
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    try:
        ins.append((r[iS].strip(), int(r[iI] or 0), int(r[iSm] or 0), [int(r[j] or 0) for _, j in stall_cols]))
    except ValueError:
        pass
tot, ts = sum(i[1] for i in ins), sum(i[2] for i in ins) or 1
print(f"{len(ins)} SASS instructions, {tot / 1e6:.1f}M warp-instructions executed, {ts} samples")
runs, cur = [], []
for k, (src, i, s, st) in enumerate(ins):
    if cur and abs(i - cur[-1][1]) > 0.15 * max(i, cur[-1][1], 1):
        runs.append(cur); cur = []
    cur.append((k, i, s, src, st))
runs.append(cur)
for run in runs:
    t = sum(x[1] for x in run); sm = sum(x[2] for x in run)
    if 100.0 * t / tot < min_share and 100.0 * sm / ts < min_share:
        continue
    ops = collections.Counter()
    for x in run:
        toks = x[3].split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        ops[op.split('.')[0]] += 1
    stl = collections.Counter()
    for x in run:
        for (name, _), v in zip(stall_cols, x[4]):
            stl[name.replace("stall_", "")] += v
    print(f"idx {run[0][0]:5d}-{run[-1][0]:5d} len {len(run):4d} exec {run[0][1] / 1e6:8.3f}M inst {100 * t / tot:5.1f}% samp {100 * sm / ts:5.1f}% | "
          + " ".join(f"{o}:{c}" for o, c in ops.most_common(6)) + " | " + " ".join(f"{n}:{100 * v / ts:.1f}" for n, v in stl.most_common(4)))
