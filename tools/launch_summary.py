"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/launch_summary.py launches.csv [skip_first_n_launches]"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("<unnamed>::", "").replace("void ", "")
    t = float(r[-1].replace(",", "")) / 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
for k, (n, t) in agg.items():
    print(f"{k:55s} n={n:4d} total_ms={t:9.4f} mean_ms={t / n:9.4f} share={t / tot:.4f}")
print(f"{'TOTAL':55s} {tot:.3f} ms")
