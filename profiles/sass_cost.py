#!/usr/bin/env python3
"""Issue-time estimate of the Newton loop of trpl_sim_kernel from its SASS, using the measured
B200 costs (profiles/r01_microbench.txt): FP64 2 cycles, DFMA with three distinct register
operands 3 cycles (2.2 with a .reuse hit), every other instruction ~0.85 cycle.
usage: sass_cost.py all.sass <function substring> [loop index from the end, default picks ~900-instr loop]"""
import re, sys, collections
path, pat = sys.argv[1], sys.argv[2]
txt = open(path).read().split("Function : ")
body = [t for t in txt if pat in t.split("\n")[0]][0]
ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", body)]
def op(s): return re.sub(r"^@!?U?P\w+\s+", "", s).split()[0].split(".")[0]
loops = []
for a, s in ins:
    if op(s) == "BRA":
        m = re.search(r"0x([0-9a-f]+)", s)
        if m and int(m.group(1), 16) <= a: loops.append((int(m.group(1), 16), a))
# Newton loop = the loop with the most MUFU per instruction among loops of 500..1500 instrs
cands = []
for lo, hi in loops:
    body_i = [s for a, s in ins if lo <= a <= hi]
    if 500 <= len(body_i) <= 1500: cands.append((lo, hi, body_i))
lo, hi, L = sorted(cands, key=lambda c: len(c[2]))[0]
c = collections.Counter(op(s) for s in L)
cyc = 0.0; n3 = 0; n3r = 0
for s in L:
    o = op(s)
    if o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        srcs = [x.strip().lstrip("-|").rstrip("|") for x in s.split(o, 1)[1].split(",")[1:]]
        regs = [x.split(".")[0] for x in srcs if re.match(r"R\d+", x)]
        if o == "DFMA" and len(set(regs)) == 3:
            if any("reuse" in x for x in srcs): cyc += 2.2; n3r += 1
            else: cyc += 3.0; n3 += 1
        else: cyc += 2.0
    else: cyc += 0.85
fp64 = c["DFMA"] + c["DMUL"] + c["DADD"]
print("loop 0x%x-0x%x: %d instr | FP64 %d (DFMA %d [3-reg %d, 3-reg+reuse %d] DMUL %d DADD %d) | SHFL %d MUFU %d FSEL %d IMAD %d other %d | est. %.0f cycles/iter"
      % (lo, hi, len(L), fp64, c["DFMA"], n3, n3r, c["DMUL"], c["DADD"], c["SHFL"], c["MUFU"], c["FSEL"], c["IMAD"],
         len(L) - fp64 - c["SHFL"] - c["MUFU"] - c["FSEL"] - c["IMAD"], cyc))
