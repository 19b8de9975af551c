"""Summarise an `ncu --page source --csv` export: stall samples and executed instructions per opcode and per
code region.  Usage: python tools/ncu_src_summary.py src.csv [chunk]"""
import csv
import sys
from collections import Counter


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


def opcode(src):
    t = src.split()
    if not t:
        return "?"
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    return op.split(".")[0]


rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 250
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ia, ie, isrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(num(r[ia]) for r in data)
tot_i = sum(num(r[ie]) for r in data)
print("kernel", rows[0][1], "| stall samples", tot, "| warp instructions", tot_i, "| SASS lines", len(data))
cs, ci = Counter(), Counter()
for r in data:
    op = opcode(r[isrc])
    cs[op] += num(r[ia])
    ci[op] += num(r[ie])
for op, c in cs.most_common(24):
    print(f"{op:10s} samples {100 * c / tot:5.1f}%   inst {100 * ci[op] / tot_i:5.1f}%")
print()
for i in range(0, len(data), chunk):
    seg = data[i:i + chunk]
    s = sum(num(r[ia]) for r in seg)
    n = sum(num(r[ie]) for r in seg)
    ops = Counter(opcode(r[isrc]) for r in seg)
    print(f"{i:5d} samples {100 * s / tot:5.1f}% inst {100 * n / tot_i:5.1f}%", ops.most_common(6))
