"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (diagnostic)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    agg.setdefault(r[ki][:70], []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"total {tot/1000:.2f} ms over {sum(len(v) for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} n={len(v):4d} mean={sum(v)/len(v):9.1f}us total={sum(v)/1000:8.2f}ms share={sum(v)/tot:6.1%}")
