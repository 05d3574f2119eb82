"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv). usage: launch_summary.py file.csv [divisor]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr, data = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    name = re.sub(r"\(.*", "", d["Kernel Name"])[:64]
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1000 if d["Metric Unit"] == "ns" else v * 1000 if d["Metric Unit"] == "ms" else v
    agg.setdefault(name, []).append(v)
tot = 0.0
for k, v in agg.items():
    print(f"{k:64s} n={len(v):5d} mean={sum(v) / len(v):9.2f} us  sum/div={sum(v) / div:10.2f}")
    tot += sum(v) / div
print(f"total/div = {tot:.1f} us")
