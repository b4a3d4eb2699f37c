#!/usr/bin/env python3
"""Per-opcode executed-instruction mix and stall samples from `ncu --page source --csv` of a report with
--import-source on.  usage: ncu_opmix.py rep.ncu-rep [frames_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter()
st = collections.Counter()
wf = collections.Counter()
wfx = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    toks = src.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.rstrip(";")
    base = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG")) else op.split(".")[0]
    n = int(r[col["Instructions Executed"]] or 0)
    ex[base] += n
    st[base] += int(r[col["# Samples"]] or 0)
    wf[base] += int(r[col["L1 Wavefronts Shared"]] or 0)
    wfx[base] += int(r[col["L1 Wavefronts Shared Excessive"]] or 0)
tot = sum(ex.values())
tots = sum(st.values())
print(f"{'opcode':14s} {'warp-instr':>12s} {'share':>7s} {'per frame':>10s} {'stall samples':>14s} {'share':>7s} {'smem wavefronts':>16s} {'excess':>9s}")
for op, n in ex.most_common(40):
    pf = f"{n / frames:10.1f}" if frames else ""
    print(f"{op:14s} {n:12d} {100 * n / tot:6.1f}% {pf} {st[op]:14d} {100 * st[op] / max(tots, 1):6.1f}% {wf[op]:16d} {wfx[op]:9d}")
print(f"{'total':14s} {tot:12d} {'':7s} {tot / frames if frames else 0:10.1f} {tots:14d}")
