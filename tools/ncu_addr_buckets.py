"""Buckets the per-instruction counts of an `ncu --page source --csv` dump by SASS address range
(diagnostic: where a kernel's executed instructions and stall samples sit)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
step = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x200
hdr = next(r for r in rows if r and r[0] == 'Address')
ia, isrc, iex, isamp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = []
for r in rows:
    if len(r) != len(hdr) or r[0] == 'Address':
        continue
    data.append((int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[isamp])))
base = data[0][0]
tot = sum(d[2] for d in data); tots = sum(d[3] for d in data)
print('total inst', tot, 'samples', tots, 'warps', data[0][2])
b = collections.OrderedDict()
for a, s, ex, sm in data:
    k = (a - base) // step * step
    v = b.setdefault(k, [0, 0]); v[0] += ex; v[1] += sm
for k, v in b.items():
    if v[0] / tot > 0.004 or v[1] / tots > 0.004:
        print(hex(k), f'inst {v[0]/tot:6.1%}  per-warp {v[0]/data[0][2]:7.1f}  samples {v[1]/tots:6.1%}')
if len(sys.argv) > 3:  # top instructions by samples
    for a, s, ex, sm in sorted(data, key=lambda d: -d[3])[:int(sys.argv[3])]:
        print(hex(a - base), f'{sm/tots:6.1%} {ex/data[0][2]:6.2f}', s[:90])
