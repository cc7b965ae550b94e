import csv, subprocess, sys, collections
rep=sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
fn=None; hdr=None
stats=collections.defaultdict(lambda: collections.Counter())
for rec in csv.reader(out.splitlines()):
    if not rec: continue
    if rec[0]=="Function Name": fn=rec[1][:60]; continue
    if rec[0]=="Address": hdr=rec; continue
    if hdr and len(rec)==len(hdr):
        d=dict(zip(hdr,rec))
        try:
            i=float(d["Instructions Executed"]); t=float(d["Thread Instructions Executed"]); s=float(d["# Samples"])
        except: continue
        if i==0: continue
        a=t/i
        b = "<6" if a<6 else "<12" if a<12 else "<20" if a<20 else "<26" if a<26 else ">=26"
        stats[fn][("inst",b)]+=i; stats[fn][("smp",b)]+=s; stats[fn][("inst","all")]+=i; stats[fn][("smp","all")]+=s
for fn,c in stats.items():
    print(fn)
    for b in ["<6","<12","<20","<26",">=26"]:
        print(f"   thr {b:5s} inst {100*c[('inst',b)]/c[('inst','all')]:5.1f}%  samples {100*c[('smp',b)]/max(c[('smp','all')],1):5.1f}%")
