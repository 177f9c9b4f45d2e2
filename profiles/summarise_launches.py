#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python profiles/summarise_launches.py gpurun_out/<tag>_launches.csv [> profiles/<tag>_launches.md]"""
import csv
import re
import sys
from collections import OrderedDict

import gzip, io
_f = io.TextIOWrapper(gzip.open(sys.argv[1])) if sys.argv[1].endswith(".gz") else open(sys.argv[1])
rows = [r for r in csv.reader(_f) if len(r) > 10 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"^void\s+", "", name)
    a = agg.setdefault(name, [0, 0.0, 1e30, 0.0, r[8], r[7]])
    t = float(r[-1]) / 1e3  # ns -> us
    a[0] += 1
    a[1] += t
    a[2] = min(a[2], t)
    a[3] = max(a[3], t)
total = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share | min us | max us | mean us |")
print("|---|---:|---:|---:|---:|---:|---:|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% | %.1f | %.1f | %.1f |" % (name, a[0], a[1] / 1e3, 100 * a[1] / total, a[2], a[3], a[1] / a[0]))
print("| **all** | %d | %.3f | 100%% | | | |" % (len(rows), total / 1e3))
