"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launches per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if d.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        k = d["Kernel Name"][:90]
        agg[k][0] += 1
        agg[k][1] += v
tot = sum(v[1] for v in agg.values())
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(f"{'us':>10} {'share':>6} {'n':>5}  kernel")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:n]:
    print(f"{v[1] / 1e3:10.1f} {v[1] / tot * 100:5.1f}% {v[0]:5d}  {k}")
print(f"{tot / 1e3:10.1f} total, {sum(v[0] for v in agg.values())} launches")
