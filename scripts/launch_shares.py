"""Per-kernel share of the step from an ncu launch list (gpu__time_duration.sum CSV).  usage: launch_shares.py <csv> [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0]
    name = name.replace("void ", "").replace("<unnamed>::", "")[-70:]
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] == "ns" else (v * 1e3 if r[iu] == "ms" else v)
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'total us':>10s} {'share':>6s} {'count':>6s} {'avg us':>9s}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t:10.1f} {100 * t / tot:5.1f}% {n:6d} {t / n:9.1f}  {k}")
