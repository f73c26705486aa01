#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: ncu_launches.py FILE.csv [skip_first_n]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[h + 1 + skip:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(',', ''))
    if r[mu] == 'us': v *= 1e3
    if r[mu] == 'ms': v *= 1e6
    name = r[kn].split('(')[0][:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s n=%4d total=%10.3f ms avg=%10.1f us %5.1f%%" % (k, v[0], v[1] / 1e6, v[1] / v[0] / 1e3, 100 * v[1] / tot))
