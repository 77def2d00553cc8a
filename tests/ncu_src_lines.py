"""Aggregates `ncu --page source --csv --print-source sass,cuda` output by CUDA source line."""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
# find header row
agg = collections.defaultdict(lambda: [0, 0, 0])
cur_file = None
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        ite = hdr.index("Thread Instructions Executed")
        ss = hdr.index("# Samples")
        continue
    if hdr and len(r) == len(hdr):
        try:
            agg[(cur_file, int(r[0]), r[1].strip()[:90])][0] += int(r[ie])
            agg[(cur_file, int(r[0]), r[1].strip()[:90])][1] += int(r[ite])
            agg[(cur_file, int(r[0]), r[1].strip()[:90])][2] += int(r[ss])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
tots = sum(v[2] for v in agg.values())
print("total warp instr", tot, "samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{k[0]:18s} {k[1]:4d} inst={v[0]/tot:6.1%} thr/inst={v[1]/max(v[0],1):5.1f} samples={v[2]/max(tots,1):6.1%}  {k[2]}")
