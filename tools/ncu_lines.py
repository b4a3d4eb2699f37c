#!/usr/bin/env python3
"""Attribute executed warp instructions of an ncu report to CUDA source lines: joins `ncu --page source --csv` (per-SASS
counts, in program order) with `nvdisasm --print-line-info` of the locally built cubin (same build!).
usage: ncu_lines.py rep.ncu-rep object.o mangled_kernel_name [frames]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, fn = sys.argv[1:4]
frames = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
counts = [(r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0),
           int(r[col["L1 Wavefronts Shared"]] or 0), int(r[col["L1 Wavefronts Shared Excessive"]] or 0)) for r in rows[2:] if len(r) >= len(hdr)]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
sec = dis[dis.index(".text." + fn + ":"):]
m = re.search(r"\n\s*\.section|\n//-+ \.text\.", sec[10:])
sec = sec[:m.start() + 10] if m else sec
line = ("?", 0); order = []
for ln in sec.splitlines():
    mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if mm:
        line = (os.path.basename(mm.group(1)), int(mm.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", ln):
        order.append(line)
assert len(order) == len(counts), (len(order), len(counts))
by = collections.Counter(); st = collections.Counter(); wf = collections.Counter(); wx = collections.Counter()
for (f_l, (src, n, s, w, x)) in zip(order, counts):
    by[f_l] += n; st[f_l] += s; wf[f_l] += w; wx[f_l] += x
tot = sum(by.values()); tots = sum(st.values())
print(f"total warp-instr {tot}  per frame {tot / frames:.0f}")
for (f, l), n in by.most_common(45):
    print(f"{f}:{l:<5d} {n / frames:10.1f}/frame {100 * n / tot:5.1f}%   stall samples {100 * st[(f, l)] / max(tots, 1):5.1f}%"
          f"   smem wavefronts {wf[(f, l)] / frames:9.1f}/frame (excess {wx[(f, l)] / frames:8.1f})")
