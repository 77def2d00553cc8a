cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_v13.json 2> gpurun_out/r02_b_v13.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_v13.json')); print('v13', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases'].items() if isinstance(v,float)})
P
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
