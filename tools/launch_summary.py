"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def summarize(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    out = ["| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append(f"| {k[:56]} | {cnt[k]} | {v / 1e3:.3f} | {v / s * 100:.1f}% | {v / cnt[k]:.1f} |")
    out.append(f"\nSum {s / 1e3:.2f} ms over {sum(cnt.values())} launches.")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
