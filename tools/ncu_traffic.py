#!/usr/bin/env python3
"""Regenerates profiles/ncu_traffic.json from this round's ncu captures (no hand-copied numbers).

Captures (tools/gpu_capture.sh, one B200, the bench.py frame: VeachAjar 1920x1080 ReSTIR PT):
    ncu --set full --cache-control none --clock-control none -k regex:traceQueue -s 39 -c 13 ...   -> one frame's 13 traversal launches
    ncu --set full --cache-control none --clock-control none -k regex:grisBounceKernel -s 18 -c 6 ...  -> one frame's 6 bounce launches
(--cache-control none: the scene's BVH stays in L2 between launches as in the real frame; ncu's default flush would show the
cold-cache traffic of a 23 MB working set instead.)

Per bench.py kernel entry the JSON holds, for the LARGEST launch (longest duration): DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum), L2 bytes (lts__t_sectors.sum x 32), duration, issue-slot utilisation, active lanes per instruction, registers,
achieved occupancy; and the per-frame sums over all launches of the entry.

Usage: ncu_traffic.py <trace.ncu-rep> [<bounce.ncu-rep>] [-o profiles/ncu_traffic.json]"""
import csv
import json
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "lts__t_sectors.sum": "lts_sectors",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
}
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"name": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in METRICS and r[i] != "":
                v = float(r[i].replace(",", ""))
                d[METRICS[h]] = v * UNIT_SCALE.get(units[i], 1.0)
        d["lts_bytes"] = d.get("lts_sectors", 0.0) * 32.0   # lts__t_sectors.sum: 32-byte sectors through the L2 tag stage
        res.append(d)
    return res


def entry(ls, source):
    big = max(ls, key=lambda d: d["duration"])
    return {
        "dram_bytes_per_launch": int(big["dram_read"] + big["dram_write"]), "lts_bytes_per_launch": int(big["lts_bytes"]),
        "duration_us": round(big["duration"], 2), "issue_active_pct": round(big["issue_active_pct"], 2),
        "threads_per_inst": round(big["threads_per_inst"], 2), "registers": int(big["registers"]),
        "achieved_occupancy_pct": round(big["achieved_occupancy_pct"], 1), "l1_hit_pct": round(big.get("l1_hit_pct", 0), 1),
        "l2_hit_pct": round(big.get("l2_hit_pct", 0), 1), "dram_pct": round(big.get("dram_pct", 0), 2), "l2_pct": round(big.get("l2_pct", 0), 2),
        "warp_instructions": int(big["warp_instructions"]), "launch": big["name"][:60],
        "frame_launches": len(ls), "frame_dram_bytes": int(sum(d["dram_read"] + d["dram_write"] for d in ls)),
        "frame_lts_bytes": int(sum(d["lts_bytes"] for d in ls)), "frame_duration_us": round(sum(d["duration"] for d in ls), 1),
        "source": source,
    }


def main():
    args = [a for a in sys.argv[1:] if a != "-o"]
    out_path = "profiles/ncu_traffic.json"
    if "-o" in sys.argv:
        out_path = sys.argv[sys.argv.index("-o") + 1]
        args.remove(out_path)
    doc = {"_comment": "written by tools/ncu_traffic.py from the .ncu-rep captures named in 'source' (ncu --set full --cache-control none "
                       "--clock-control none, bench.py frame, one B200); figures of the LARGEST launch of each entry + per-frame sums"}
    tq = launches(args[0])
    src = f"{args[0]} ({len(tq)} launches of one frame)"
    closest = [d for d in tq if "<0>" in d["name"] or "(int)0" in d["name"] or "ILi0" in d["name"]]
    anyhit = [d for d in tq if d not in closest]
    reuse = anyhit[-2:] if len(anyhit) > 2 else anyhit
    path_any = anyhit[:-2] if len(anyhit) > 2 else []
    if not closest:
        closest = anyhit
    doc["trace_closest"] = entry(closest, src)
    doc["trace_any"] = entry(reuse, src + ", the two visibility launches of the reuse passes")
    doc["trace_paths"] = entry(closest + path_any, src + ", the path tracer's closest-hit and any-hit launches")
    if len(args) > 1:
        doc["gris_bounce"] = entry(launches(args[1]), f"{args[1]}")
    json.dump(doc, open(out_path, "w"), indent=1)
    for k, v in doc.items():
        if isinstance(v, dict):
            print(f"{k:14s} {v['duration_us']:8.1f} us  dram {v['dram_bytes_per_launch']/1e6:8.1f} MB  lts {v['lts_bytes_per_launch']/1e6:8.1f} MB  "
                  f"issue {v['issue_active_pct']:5.1f} %  lanes {v['threads_per_inst']:5.2f}  regs {v['registers']}  occ {v['achieved_occupancy_pct']} %")


if __name__ == "__main__":
    main()
