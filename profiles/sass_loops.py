#!/usr/bin/env python3
"""List backward branches (loops) of one kernel in a cuobjdump -sass dump with the per-loop
instruction mix.  usage: sass_loops.py all.sass <substring of function name>"""
import re, sys, collections
path, pat = sys.argv[1], sys.argv[2]
txt = open(path).read().split("Function : ")
body = [t for t in txt if pat in t.split("\n")[0]][0]
ins = []
for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", body):
    ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
def op(s):
    s = re.sub(r"^@!?U?P\w+\s+", "", s)
    return s.split()[0].split(".")[0]
loops = []
for a, s in ins:
    if op(s) == "BRA":
        m = re.search(r"0x([0-9a-f]+)", s)
        if m and int(m.group(1), 16) <= a:
            loops.append((int(m.group(1), 16), a))
for lo, hi in sorted(loops, key=lambda x: x[1]-x[0]):
    c = collections.Counter(op(s) for a, s in ins if lo <= a <= hi)
    n = sum(c.values())
    fp64 = c["DFMA"] + c["DMUL"] + c["DADD"]
    print("loop 0x%x-0x%x: %d instr, FP64 %d (DFMA %d DMUL %d DADD %d) MUFU %d SHFL %d LDS %d STS %d other %d" % (
        lo, hi, n, fp64, c["DFMA"], c["DMUL"], c["DADD"], c["MUFU"], c["SHFL"], c["LDS"], c["STS"],
        n - fp64 - c["MUFU"] - c["SHFL"] - c["LDS"] - c["STS"]))
    if "-v" in sys.argv:
        print("   ", dict(c.most_common(25)))
