"""Stall-reason histogram per SASS index range of an `ncu --page source --csv` export.
Usage: ncu_region_stalls.py src.csv lo:hi [lo:hi ...]   (indices as printed by ncu_src_summary.py; the export
repeats the kernel twice, use the first half)"""
import csv
import sys


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
half = len(data) // 2
if [r[1] for r in data[:half]] == [r[1] for r in data[half:2 * half]]:
    data = data[:half]
ia, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(num(r[ia]) for r in data)
print("SASS lines", len(data), "samples", tot)
for rng in sys.argv[2:]:
    lo, hi = (int(x) for x in rng.split(":"))
    seg = data[lo:hi]
    s = sum(num(r[ia]) for r in seg)
    n = sum(num(r[ie]) for r in seg)
    hist = sorted(((sum(num(r[j]) for r in seg), h.replace("stall_", "")) for j, h in reasons), reverse=True)
    print(f"[{lo}:{hi}] samples {100 * s / tot:5.1f}%  inst {n}  " +
          ", ".join(f"{h} {100 * c / max(s, 1):.0f}%" for c, h in hist[:8] if c))
