#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [> profiles/rNN_launches.md]"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        short = re.sub(r"\(.*", "", r[ki])
        short = re.sub(r"^.*::", "", short)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, a[0], a[1], 100 * a[1] / tot))
    print("\n%d launches, %.1f us total (cold-cache, serialised: compare shares, not absolutes)" % (n, tot))


if __name__ == "__main__":
    main(sys.argv[1])
