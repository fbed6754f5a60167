#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (lambda owner + index)."""
import collections
import csv
import re
import sys


def main(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    for r in rows:
        n = r["Kernel Name"]
        m = re.search(r"ZbPipe::(\w+)\([^)]*\)::\{lambda\(long\)#(\d+)\}", n)
        key = "%s#%s" % (m.group(1), m.group(2)) if m else n.split("(")[0].replace("void ", "")[:48]
        t = float(r["Metric Value"]) / 1e6
        a = agg.setdefault(key, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print("%-44s %6s %10s %7s  %s" % ("kernel", "n", "ms", "share", "grid x block (first launch)"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-44s %6d %10.3f %7.3f  %s x %s" % (k, a[0], a[1], a[1] / tot, a[2], a[3]))
    print("total %.3f ms over %d launches" % (tot, len(rows)))


if __name__ == "__main__":
    main(sys.argv[1])
