"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) by kernel:
launches, total ms, share, average us -- as a markdown table.  Usage: python tools/launch_summary.py X.csv [skip_regex]"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:60]


def main():
    path = sys.argv[1]
    skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = [line for line in open(path, newline="") if line.startswith('"')]
    rd = csv.DictReader(rows)
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        if skip and skip.search(k):
            continue
        tot[k] += float(r["Metric Value"].replace(",", "")) * (1e-6 if r["Metric Unit"] == "ns" else 1e-3)
        cnt[k] += 1
    total = sum(tot.values())
    print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
    for k in sorted(tot, key=lambda x: -tot[x]):
        print(f"| {k} | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / total:.1f}% | {1e3 * tot[k] / cnt[k]:.1f} |")
    print(f"\nSum {total:.2f} ms over {sum(cnt.values())} launches.")


if __name__ == "__main__":
    main()
