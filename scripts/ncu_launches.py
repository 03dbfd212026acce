"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ik, iv, im, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1)
    a = agg.setdefault(r[ik][:100], [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"{len(data)} launches, total {tot / 1e6:.3f} ms (cold-cache, serialised: compare SHARES, not absolutes)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"{t / 1e6:12.3f} ms {100 * t / tot:6.2f}%  x{n:<4d} {k}")
