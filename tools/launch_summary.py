"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    key = (re.sub(r"\(.*", "", row["Kernel Name"])[:70], row["Grid Size"])
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"{'us':>10} {'n':>4} {'avg us':>9} {'share':>6}  kernel (grid)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {n:4d} {t/n:9.1f} {100*t/tot:5.1f}%  {k[0]} {k[1]}")
print(f"total {tot:.1f} us over {sum(n for n,_ in agg.values())} launches")
