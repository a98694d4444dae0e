"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.  usage: launch_shares.py launches.csv [skip_first_n]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
n = 0
for r in rows:
    if r is hdr or r[mi] != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    name = r[ki][:90]
    t, c = agg.get(name, (0.0, 0))
    agg[name] = (t + float(r[vi].replace(",", "")), c + 1)
tot = sum(t for t, _ in agg.values())
print(f"total {tot/1e6:.3f} ms in {sum(c for _, c in agg.values())} launches")
for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}% x{c:5d}  {name}")
