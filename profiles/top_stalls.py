"""Top warp-stall SASS lines per kernel from `ncu -i X.ncu-rep --page source --csv` (stdin)."""
import csv, sys
kernel=None; rows=[]; out=[]
def flush():
    if kernel and rows:
        tot=sum(r[1] for r in rows) or 1
        print(f"== {kernel[:100]}  (samples {tot})")
        for src,s,ex,conf in sorted(rows, key=lambda r:-r[1])[:int(sys.argv[1]) if len(sys.argv)>1 else 14]:
            print(f"  {100*s/tot:5.1f}%  exec={ex:>7}  smem_nway={conf:>4}  {src.strip()[:100]}")
for r in csv.reader(sys.stdin):
    if not r: continue
    if r[0]=="Kernel Name": flush(); kernel=r[1]; rows=[]; continue
    if r[0]=="Address": hdr=r; si=hdr.index("Warp Stall Sampling (All Samples)"); ei=hdr.index("Instructions Executed"); ci=hdr.index("L1 Conflicts Shared N-Way"); continue
    try: rows.append((r[1], float(r[si]), r[ei], r[ci]))
    except Exception: pass
flush()
