#!/usr/bin/env python3
"""Summarise an ncu launch list (gpu__time_duration.sum csv): per-kernel count, mean, share; optional per-launch dump."""
import csv, io, sys, collections
lines = open(sys.argv[1]).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(io.StringIO("\n".join(lines[i:]))))
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("rt::", "").replace("<unnamed>::", "")[:48]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
for k, v in agg.items():
    print(f"{v[0]:5d} launches  mean {v[1]/v[0]/1e3:9.1f} us  total {v[1]/1e6:8.3f} ms  share {100*v[1]/tot:5.1f}%  {k}")
print(f"total {tot/1e6:.3f} ms over {len(rows)} launches")
if len(sys.argv) > 2:
    for r in rows[: int(sys.argv[2])]:
        print(f"{float(r['Metric Value'])/1e3:9.1f} us  {r['Grid Size']:>14s} {r['Kernel Name'].split('(')[0][-40:]}")
