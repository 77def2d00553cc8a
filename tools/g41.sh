cd $GRAFT_REPO_ROOT
( time timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) 2>&1 | tail -9
for k in 1 2; do
HB2_BENCH_CPU=0 timeout 300 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_v19_$k.json 2> gpurun_out/r02_b_v19_$k.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_v19_$k.json')); print('v19', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases'].items() if isinstance(v,float)}, d['e2e']['value'])
P
done
