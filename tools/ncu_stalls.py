#!/usr/bin/env python3
"""Warp-stall reasons (sampled) per SASS instruction and in total, executed-instruction mix and shared-memory wavefronts per
frame from `ncu --page source --csv` of a report taken with --import-source on.
usage: ncu_stalls.py rep.ncu-rep [frames_per_launch] [top_n]"""
import csv,sys,subprocess,io
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[1]; col={h:i for i,h in enumerate(hdr)}
R=['stall_long_sb','stall_short_sb','stall_wait','stall_math','stall_not_selected','stall_selected','stall_mio','stall_lg','stall_dispatch','stall_branch_resolving','stall_barrier','stall_no_inst','stall_misc','stall_membar','stall_drain','stall_sleep','stall_tex']
tot={r:0 for r in R}
data=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    d={k:int(r[col[k]] or 0) for k in R}
    for k in R: tot[k]+=d[k]
    data.append((r[col['Source']].strip(), int(r[col['# Samples']] or 0), d, int(r[col['Instructions Executed']] or 0), int(r[col['L1 Wavefronts Shared']] or 0)))
S=sum(tot.values())
print({k:round(100*v/S,1) for k,v in tot.items() if v})
frames=float(sys.argv[2]) if len(sys.argv)>2 else 1
print("instr/frame", sum(d[3] for d in data)/frames, "smem wf/frame", sum(d[4] for d in data)/frames)
import collections
op=collections.Counter(); 
for src,n,d,ex,wf in data: op[src.split()[0] if not src.startswith('@') else src.split()[1]]+=ex
print([(k,round(v/frames,1)) for k,v in op.most_common(40)])
N=int(sys.argv[3]) if len(sys.argv)>3 else 30
for i,(src,n,d,ex,wf) in sorted(enumerate(data), key=lambda t:-t[1][1])[:N]:
    top=sorted(d.items(), key=lambda kv:-kv[1])[:2]
    print(i, n, round(100*n/S,1), src[:80], top)
