"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: last iteration, per kernel."""
import collections
import csv
import sys

path, marker = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "perturb_forward_kernel")
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = [(r["Kernel Name"].split("(")[0].replace("icadv::", ""), float(r["Metric Value"].replace(",", "")), r["Grid Size"]) for r in rows]
idx = [i for i, p in enumerate(per) if marker in p[0]]
last = per[idx[-1]:] if len(idx) < 2 else per[idx[-2]:idx[-1]]
tt = sum(v for _, v, _ in last)
agg = collections.OrderedDict()
for k, v, _ in last:
    agg[k] = agg.get(k, 0) + v
for k, v in agg.items():
    print(f"{k:40s} {v/1e3:10.1f} us {100*v/tt:5.1f}%")
print(f"total {tt/1e3:.1f} us over {len(last)} launches")
if "-v" in sys.argv:
    for i, (k, v, g) in enumerate(last):
        print(i, k[:28], g, round(v / 1e3, 1))
