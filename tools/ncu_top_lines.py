"""Top stalled SASS lines of an `ncu --page source --csv` export.  Usage: ncu_top_lines.py src.csv [N]"""
import csv
import sys


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ia, ie, isrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(num(r[ia]) for r in data)
top = sorted(range(len(data)), key=lambda i: -num(data[i][ia]))[:n]
for i in sorted(top):
    r = data[i]
    rs = sorted([(num(r[j]), h.replace("stall_", "")) for j, h in reasons], reverse=True)[:2]
    print(f"{i:6d} {100 * num(r[ia]) / tot:5.1f}% exec {num(r[ie]):9d}  {r[isrc][:60]:60s} {[x for x in rs if x[0]]}")
