#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/xxx.md"""
import collections
import csv
import io
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k, cnt[k], v, v / cnt[k], 100 * v / T))
    print("| **all** | %d | %.1f | | 100%% |" % (sum(cnt.values()), T))


if __name__ == "__main__":
    main(sys.argv[1])
