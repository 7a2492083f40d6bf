#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name (and per grid with -g)."""
import collections, csv, re, sys
path = sys.argv[1]
bygrid = "-g" in sys.argv
top = 40
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ts::", "").replace("ts::", "")
    key = (name, row["Grid Size"]) if bygrid else (name, "")
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for (k, g), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d}  avg={t / n:8.1f}  {k} {g}")
