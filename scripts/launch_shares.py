"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kn = h.index("Kernel Name"); mv = h.index("Metric Value"); mu = h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) > mv:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1e-3)
        n = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
        agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':72s} {'launches':>8s} {'total us':>12s} {'mean us':>10s} {'share':>7s}")
for n, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n:72s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {100 * v[1] / tot:6.1f}%")
