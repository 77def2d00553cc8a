cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 900 python -m pytest tests/test_filters.py tests/test_gpu_parity.py tests/test_writers.py tests/test_bonds.py -q -m gpu -k "not drift" 2>&1 | tail -12
timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_v16.json 2> gpurun_out/r02_b_v16.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_v16.json')); print('v16', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases'].items() if isinstance(v,float)}, d['e2e']['value'])
P
