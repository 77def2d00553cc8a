cd $GRAFT_REPO_ROOT
( time timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 ) 2>&1
HB2_BENCH_CPU=0 timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_v17.json 2> gpurun_out/r02_b_v17.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_v17.json')); print('v17', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases'].items() if isinstance(v,float)}, d['e2e']['value'])
P
