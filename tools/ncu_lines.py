#!/usr/bin/env python3
"""Summarise an .ncu-rep per CUDA source line (samples, threads/instruction).  Usage: ncu_lines.py rep [topN]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur_file = None; hdr = None; rows = []
for rec in csv.reader(out.splitlines()):
    if not rec: continue
    if rec[0] == "File Path": cur_file = rec[1].split("/")[-1]; continue
    if rec[0] == "Function Name": continue
    if rec[0] == "Line No": hdr = rec; continue
    if hdr and rec[0] != "" and len(rec) == len(hdr):
        d = {}
        for k, v in zip(hdr, rec):
            d.setdefault(k, v)
        rows.append((cur_file, d))
def num(x):
    try: return float(x)
    except: return 0.0
tot = sum(num(r["# Samples"]) for _, r in rows)
inst = sum(num(r["Instructions Executed"]) for _, r in rows)
thr = sum(num(r["Thread Instructions Executed"]) for _, r in rows)
print(f"total samples {tot:.0f}  warp-instructions {inst:.3g}  avg threads/instr {thr/max(inst,1):.2f}")
byfile = collections.Counter(); instfile = collections.Counter(); thrfile = collections.Counter()
for f, r in rows:
    byfile[f] += num(r["# Samples"]); instfile[f] += num(r["Instructions Executed"]); thrfile[f] += num(r["Thread Instructions Executed"])
for f, s in byfile.most_common():
    print(f"  {f:24s} samples {100*s/tot:5.1f}%  instr {100*instfile[f]/inst:5.1f}%  threads/instr {thrfile[f]/max(instfile[f],1):5.1f}")
rows.sort(key=lambda x: -num(x[1]["# Samples"]))
for f, r in rows[:top]:
    i = num(r["Instructions Executed"])
    print(f"{100*num(r['# Samples'])/tot:5.1f}% inst {100*i/inst:5.1f}% thr {num(r['Thread Instructions Executed'])/max(i,1):4.1f} {f}:{r['Line No']} {r['Source'].strip()[:100]}")
