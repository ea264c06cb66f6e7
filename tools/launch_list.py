#!/usr/bin/env python3
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` log into id,kernel,grid,block,us lines
plus per-kernel totals.   python tools/launch_list.py <log.csv> [last_n]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
last = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows)
tot = collections.OrderedDict()
print("id,kernel,grid,block,us")
for r in rows[-last:]:
    name = r[4].split("(")[0].replace("void ", "").replace("pymd::<unnamed>::", "").replace("pymd::", "")
    us = float(r[-1].replace(",", "")) / 1e3
    print("%s,%s,%s,%s,%.2f" % (r[0], name, r[7].replace(",", "x"), r[6].replace(",", "x"), us))
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += us
s = sum(t[1] for t in tot.values())
for k, (n, us) in tot.items():
    print("# %-40s n=%4d  total %10.1f us  share %.3f" % (k, n, us, us / s))
