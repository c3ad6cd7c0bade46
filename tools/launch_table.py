"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name: python tools/launch_table.py file.csv"""
import csv, sys, collections, re
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    us = v / 1000 if u.startswith("n") else (v * 1000 if u.startswith("m") else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]); name = re.sub(r"^void ", "", name)[:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1000:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us/1000:8.3f} ms {us/tot*100:5.1f}%  n={n:4d}  avg {us/n:7.1f} us  {k}")
