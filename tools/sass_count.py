#!/usr/bin/env python3
"""Static instruction mix of one kernel in an object file (experiments: how many instructions does a variant issue per frame).
usage: sass_count.py file.o substring-of-mangled-name [first-index last-index]"""
import collections, re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for part in txt.split("Function : ")[1:]:
    name = part.split("\n")[0]
    if sys.argv[2] not in name:
        continue
    ops = [m.group(2).strip() for m in (re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l) for l in part.split("\n")) if m]
    a, b = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, len(ops))
    # main loop = from the first STS.64 burst's preceding 600 instructions to the last STG (heuristic when no range is given)
    if len(sys.argv) <= 4:
        stg = [i for i, o in enumerate(ops) if "STG" in o]
        back = [i for i, o in enumerate(ops) if re.search(r"BRA\s+0x", o)]
        b = stg[-1] + 1
        lds = [i for i, o in enumerate(ops) if re.match(r"(@!?P\d+\s+)?LDS\s", o)]
        a = min(i for i in lds if i > 300) - 30
    c = collections.Counter((o.split()[1] if o.startswith("@") else o.split()[0]) for o in ops[a:b])
    print(name, "total", len(ops), "range", a, b, "count", b - a)
    print(c.most_common(30))
    break
