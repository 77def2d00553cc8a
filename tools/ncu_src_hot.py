"""Per-SASS-instruction execution counts of the first kernel in an ncu report: plateaus of equal count
(loop structure) and the instructions with the most stall samples. usage: ncu_src_hot.py report.ncu-rep [n_top [kernel-name substring]]"""
import csv, subprocess, sys, io
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out = []; k = 0; want = sys.argv[3] if len(sys.argv) > 3 else ""; on = False
for r in rows:
    if r and r[0] == "Kernel Name":
        if on: break
        on = want in r[1]
        if on: print(r[1][:100])
        continue
    if r and r[0] == "Address": hdr = r; continue
    if on and len(r) > 10: out.append(r)
ix = {h: i for i, h in enumerate(hdr)}
E_ = ix["Instructions Executed"]
tot = sum(int(r[E_]) for r in out)
print("total warp instructions", tot, "static", len(out))
prev = None; start = 0; acc = 0; segs = []
for i, r in enumerate(out):
    c = int(r[E_])
    if prev is None or abs(c - prev) > 0.03 * max(c, prev, 1):
        if prev is not None: segs.append((start, i - 1, prev, acc))
        start = i; acc = 0
    prev = c; acc += c
segs.append((start, len(out) - 1, prev, acc))
for s in segs:
    if s[3] > 0.004 * tot:
        print(f"{s[0]:5d}-{s[1]:5d} n={s[1]-s[0]+1:4d} count~{s[2]:9d} share={s[3]/tot*100:5.1f}%  {out[s[0]][1].strip()[:60]}")
w = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in out); wi = sum(int(r[ix["L1 Wavefronts Shared Ideal"]] or 0) for r in out)
print("smem wavefronts", w, "ideal", wi)
st = sum(int(r[2]) for r in out)
top = sorted(range(len(out)), key=lambda i: -int(out[i][2]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    r = out[i]; print(f"{i:5d} {int(r[2])/st*100:5.1f}% exec={r[E_]:>9} wf={r[ix['L1 Wavefronts Shared']]:>9} {r[1].strip()[:70]}")
