cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 40 --warmup 5 > gpurun_out/r02_launches_final.log 2>&1
python - <<'P'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_launches_final.csv')))
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
            if u=='ns': v/=1000
            elif u=='ms': v*=1000
            k=d['Kernel Name'][:70]; agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{v[1]:10.1f} us {v[0]:5d} launches {v[1]/v[0]:9.2f} us/launch {100*v[1]/tot:5.1f}%  {k}")
P
