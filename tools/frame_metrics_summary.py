#!/usr/bin/env python3
"""Sums the per-launch counters of tools/gpu_frame_metrics.sh by kernel: warp instructions, lanes per instruction, DRAM / L2 bytes, time."""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; iN, iM, iV = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
iID = h.index("ID")
per = collections.defaultdict(dict)
for r in rows[1:]:
    per[(int(r[iID]), r[iN])][r[iM]] = float(r[iV].replace(",", ""))
agg = collections.defaultdict(lambda: collections.Counter())
for (_, name), m in per.items():
    k = name.split("(")[0].replace("void rt::<unnamed>::", "").replace("void unnamed>::", "")
    a = agg[k]; a["n"] += 1
    for key, v in m.items(): a[key] += v
tot = collections.Counter()
for a in agg.values(): tot.update(a)
print(f"{'kernel':42s} {'n':>4s} {'ms':>8s} {'Mwarp-inst':>11s} {'share':>6s} {'lanes':>6s} {'DRAM MB':>9s} {'L2 MB':>9s} {'occ %':>6s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["smsp__inst_executed.sum"]):
    wi = a["smsp__inst_executed.sum"]
    print(f"{k[:42]:42s} {a['n']:4d} {a['gpu__time_duration.sum']/1e6:8.3f} {wi/1e6:11.1f} {100*wi/tot['smsp__inst_executed.sum']:5.1f}% {a['smsp__thread_inst_executed.sum']/max(wi,1):6.1f} "
          f"{a['dram__bytes.sum']/1e6:9.1f} {a['lts__t_bytes.sum']/1e6:9.1f} {a['sm__warps_active.avg.pct_of_peak_sustained_active']/a['n']:6.1f}")
wi = tot["smsp__inst_executed.sum"]
print(f"total: {tot['n']} launches, {tot['gpu__time_duration.sum']/1e6:.3f} ms serialised, {wi/1e6:.1f} M warp instructions, "
      f"{tot['smsp__thread_inst_executed.sum']/wi:.1f} lanes/instr, DRAM {tot['dram__bytes.sum']/1e9:.2f} GB, L2 {tot['lts__t_bytes.sum']/1e9:.2f} GB")
