"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_shares.py launches.csv"""
import csv
import sys
from collections import defaultdict

tot, cnt = defaultdict(float), defaultdict(int)
with open(sys.argv[1]) as f:
    rows = [r for r in csv.reader(f) if len(r) > 10 and r[0].isdigit()]
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    if name.startswith("k_sweep"):
        name = r[4].split("(MmaArgs")[0].replace("void ", "")
    tot[name] += float(r[-1]) / 1e6
    cnt[name] += 1
total = sum(tot.values())
print(f"{len(rows)} launches, {total:.3f} ms device time in total (cold-cache, serialised: compare shares)")
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:60s} {cnt[k]:8d} {tot[k]:10.3f} {100 * tot[k] / total:6.1f}%")
