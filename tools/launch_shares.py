#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel + grid: launches, total and share of time.
    python tools/launch_shares.py launches.csv [skip_first_n]"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "")
    key = (name, r[8])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += float(r[14].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot/1e6:.3f} ms summed (cold-cache, serialised)")
for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*t/tot:5.1f}%  {t/1e3:10.1f} us  x{n:<5d} avg {t/n/1e3:9.1f} us  {name} grid {grid}")
