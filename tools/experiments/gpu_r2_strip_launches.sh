#!/bin/bash
# launch list (durations) of the last frame of one standalone strip, serial mode: gpu_r2_strip_launches.sh R0 R1 tag [ENV=1 ...]
R0=$1; R1=$2; TAG=$3; shift 3
mkdir -p gpurun_out
env RPT_NO_FRAME_OVERLAP=1 "$@" timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --cache-control none --clock-control none \
   -k "regex:^(?!.*(ploc|Ploc|flatten|morton|initLeaves|collapse|DeviceRadixSort|DeviceScan))" --csv --log-file gpurun_out/$TAG.csv \
   python tools/experiments/gpu_r2_strip_alone.py one $R0 $R1 > gpurun_out/$TAG.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/$TAG.csv") if l.startswith('"'))]
h = rows[0]; iN, iM, iV, iID = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((int(r[iID]), r[iN].split("(")[0].replace("void rt::<unnamed>::", "").replace("void unnamed>::", "")), {})[r[iM]] = float(r[iV].replace(",", ""))
items = list(per.items())
# the last frame = from the last gbufferKernel on
start = max(i for i, ((_, n), _) in enumerate(items) if n.startswith("gbufferKernel"))
tot = 0
for (id_, n), m in items[start:]:
    d = m["gpu__time_duration.sum"] / 1e3; tot += d
    wi = m["smsp__inst_executed.sum"]
    print(f"{n[:40]:40s} {d:9.1f} us  {wi/1e6:8.2f} M warp-instr  {m['smsp__thread_inst_executed.sum']/max(wi,1):5.1f} lanes")
print(f"total {tot/1e3:.3f} ms serialised")
PY
