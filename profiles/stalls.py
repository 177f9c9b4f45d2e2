#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: stall reasons and hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, src = hdr.index('# Samples'), hdr.index('Source')
data = [r for r in rows[2:] if len(r) == len(hdr) and r[si].isdigit()]
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[si]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {}
for r in data:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print('stall totals:', [(k, v, round(v / max(1, sum(agg.values())), 3)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
    st = sorted([(hdr[i], int(r[i] or 0)) for i in stalls], key=lambda kv: -kv[1])[:2]
    print('%6s %.3f  %-72s %s' % (r[si], int(r[si]) / tot, r[src].strip()[:72], st))
