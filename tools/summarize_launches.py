#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table
(markdown) for profiles/.  Usage: tools/summarize_launches.py launches.csv "title" > profiles/x.md"""
import collections
import csv
import re
import sys


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "launch list")
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0, ""])
    tot = 0.0
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit == "ns" else v / 1e3 if unit.startswith("us") else v
        a = agg[name]
        a[0] += 1; a[1] += v; a[2] = f'{row["Grid Size"]} x {row["Block Size"]}'
        tot += v
    print(f"# {title}\n")
    print("Per-launch device times are cold-cache and serialised under ncu: compare SHARES, not absolutes.\n")
    print(f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
    print("| kernel | launches | total ms | mean us | share | last grid x block |")
    print("|---|---:|---:|---:|---:|---|")
    for k, (n, t, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.3f} | {1e3 * t / n:.1f} | {t / tot:.3f} | {g} |")


if __name__ == "__main__":
    main()
