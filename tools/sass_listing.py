#!/usr/bin/env python
"""ncu `--page source --csv` export -> compact per-instruction listing: index, executions per unit, stall samples, SASS.
    python tools/sass_listing.py src.csv <units e.g. warps*rows> > listing.txt"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
seen, data = set(), []
for r in rows[hi + 1:]:
    if r and r[0].startswith('0x') and r[0] not in seen:
        seen.add(r[0]); data.append(r)
ie, ss, src = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
tot = sum(int(r[ie]) for r in data)
print('# instructions %d, executed %d (%.1f per unit), samples %d' % (len(data), tot, tot / units, sum(int(r[ss]) for r in data)))
for i, r in enumerate(data):
    print('%4d %6.2f %6s  %s' % (i, int(r[ie]) / units, r[ss], r[src][:110]))
