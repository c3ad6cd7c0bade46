import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; ni, mi, ui, vi, ii = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value'), h.index('ID')
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
t = defaultdict(float); d = defaultdict(float); nm = {}
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', ''))
    if r[mi].startswith('gpu__time'): t[r[ii]] += v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1)
    elif r[mi].startswith('dram__bytes'): d[r[ii]] += v * U.get(r[ui], 1)
    nm[r[ii]] = r[ni].split('(')[0].replace('void ', '')[:48]
agg = defaultdict(lambda: [0, 0.0, 0.0])
for k in t: a = agg[nm[k]]; a[0] += 1; a[1] += t[k]; a[2] += d[k]
tot = sum(a[1] for a in agg.values())
print(f"# one eager forward under ncu (serialized, cold cache): {len(t)} launches, {tot/1e3:.3f} ms")
print(f"# {'time_us':>10} {'n':>4} {'share':>6} {'dram_GB':>8} {'GB/s':>7}  kernel")
for n, (c, us, by) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {us:10.1f} {c:4d} {us/tot*100:5.1f}% {by/1e9:8.2f} {by/us/1e3 if us else 0:7.0f}  {n}")
