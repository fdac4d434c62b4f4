#!/usr/bin/env python3
"""Aggregate the ncu source page (--page source --csv, SASS view) by opcode and list the
instructions with the most stall samples."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
byop = collections.Counter(); stall = collections.defaultdict(collections.Counter); execd = collections.Counter()
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    src = r[ix["Source"]]
    op = re.sub(r"^@!?U?P\w+\s+", "", src).split()[0].split(".")[0] if src else "?"
    n = int(r[ix["# Samples"]] or 0)
    byop[op] += n
    execd[op] += int(r[ix["Instructions Executed"]] or 0)
    for h in reasons:
        stall[op][h] += int(r[ix[h]] or 0)
print("total samples", tot)
print("%-8s %8s %6s %12s  top stall reasons" % ("op", "samples", "%", "executed"))
for op, n in byop.most_common(14):
    top = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in stall[op].most_common(4))
    print("%-8s %8d %5.1f%% %12d  %s" % (op, n, 100.0 * n / tot, execd[op], top))
print("\nhottest instructions:")
hot = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in hot:
    rs = sorted(((int(r[ix[h]] or 0), h.replace("stall_", "")) for h in reasons), reverse=True)[:3]
    print("%s %6s  %-60s %s" % (r[ix["Address"]][-5:], r[ix["# Samples"]], r[ix["Source"]][:60], rs))
