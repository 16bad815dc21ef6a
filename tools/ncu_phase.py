"""Aggregate the stall samples of an ncu source page (cuda,sass) by source-line ranges.
Usage: python tools/ncu_phase.py mix.csv name:file:lo-hi ..."""
import csv
import sys

def main():
    ranges = []
    for a in sys.argv[2:]:
        name, f, r = a.split(":")
        lo, hi = r.split("-")
        ranges.append((name, f, int(lo), int(hi)))
    cur = None
    hdr = None
    samp = {}
    inst = {}
    other = {}
    for r in csv.reader(open(sys.argv[1])):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 8 or r[0] == "":
            continue
        try:
            ln, s, i = int(r[0]), int(r[4]), int(r[7])
        except ValueError:
            continue
        for name, f, lo, hi in ranges:
            if f == cur and lo <= ln <= hi:
                samp[name] = samp.get(name, 0) + s
                inst[name] = inst.get(name, 0) + i
                break
        else:
            other[(cur, ln)] = other.get((cur, ln), 0) + s
            samp["(other)"] = samp.get("(other)", 0) + s
            inst["(other)"] = inst.get("(other)", 0) + i
    tot = sum(samp.values())
    ti = sum(inst.values())
    for k in sorted(samp, key=lambda x: -samp[x]):
        print(f"{k:24s} samples {100 * samp[k] / tot:5.1f}%   warp-instructions {100 * inst[k] / ti:5.1f}%")
    print("largest unassigned:", sorted(other.items(), key=lambda x: -x[1])[:8])

if __name__ == "__main__":
    main()
