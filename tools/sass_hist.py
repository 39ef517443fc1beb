#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel:  tools/sass_hist.py <object or .so> <substring of the mangled name> [--loop]
--loop: only the instructions between the first backward-branch target and the last backward branch (the main loop)."""
import re
import subprocess
import sys
from collections import Counter

obj, pat = sys.argv[1], sys.argv[2]
loop = "--loop" in sys.argv
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
start = None
for n, l in enumerate(out):
    if "Function :" in l:
        if start is not None:
            end = n
            break
        if pat in l:
            start = n
else:
    end = len(out)
ins = []
rx = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
for l in out[start:end]:
    m = rx.match(l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
if loop:
    tgt = [(a, int(re.search(r"0x([0-9a-f]+)", t).group(1), 16)) for a, t in ins if re.search(r"\bBRA\b", t) and re.search(r"0x[0-9a-f]+", t)]
    back = [(a, t) for a, t in tgt if t < a]
    lo = min(t for a, t in back); hi = max(a for a, t in back)
    ins = [(a, t) for a, t in ins if lo <= a <= hi]
    print(f"loop 0x{lo:x}..0x{hi:x}")
c = Counter()
for a, t in ins:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t.split()[0]
    c[op.split(".")[0]] += 1
tot = sum(c.values())
print("instructions", tot)
fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("FP64-pipe (DFMA DMUL DADD DSETP DMNMX):", fp64, f"{100*fp64/tot:.1f}%")
for k, v in c.most_common(40):
    print(f"  {k:12s} {v:6d} {100*v/tot:5.1f}%")
