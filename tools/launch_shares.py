"""Per-kernel shares of one forward from an ncu launch list csv."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; ni = h.index('Kernel Name'); vi = h.index('Metric Value')
recs = [(r[ni], float(r[vi].replace(',', ''))) for r in rows[hi + 1:] if len(r) > vi]
agg = defaultdict(lambda: [0, 0.0])
for n, v in recs:
    k = n.split('(')[0].replace('void ', '')[:60]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v for _, v in recs)
print(f"launches {len(recs)}  total {tot/1e6:.3f} ms (ncu: serialized, cold cache)")
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:20]:
    print(f"{v/1e3:10.1f} us {c:4d} {v/tot*100:5.1f}%  {n}")
if len(sys.argv) > 2:
    for n, v in recs: print(f"{v/1e3:9.1f}  {n[:110]}")
