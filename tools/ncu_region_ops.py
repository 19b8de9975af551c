"""Stall samples of an `ncu --page source --csv` export aggregated by opcode inside instruction-index regions.
Usage: ncu_region_ops.py src.csv split1[,split2...]   (regions: [0,split1), [split1,split2), ...)"""
import csv
import sys
from collections import defaultdict


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
splits = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else []
bounds = [0] + splits + [len(data)]
ia, ie, isrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
reasons = [(i, h.replace("stall_", "")) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(num(r[ia]) for r in data)
for lo, hi in zip(bounds[:-1], bounds[1:]):
    ops, why, execs = defaultdict(int), defaultdict(int), defaultdict(int)
    for r in data[lo:hi]:
        op = r[isrc].split()
        op = (op[1] if op and op[0].startswith("@") else op[0]) if op else "?"
        op = op.split(".")[0]
        ops[op] += num(r[ia])
        execs[op] += num(r[ie])
        for j, h in reasons:
            why[h] += num(r[j])
    s = sum(ops.values())
    print(f"region [{lo},{hi}): {100 * s / tot:.1f}% of samples")
    print("   by op:", ", ".join(f"{k} {100 * v / tot:.1f}% (x{execs[k]})" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:10]))
    print("   by reason:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(why.items(), key=lambda kv: -kv[1])[:8]))
