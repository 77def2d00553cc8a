cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
for coop in 1 0 1 0; do
HB2_FUSE_COOP=$coop timeout 300 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_coop$coop.json 2> gpurun_out/r02_b_coop$coop.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_coop$coop.json')); print('coop=$coop', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases'].items() if isinstance(v,float)}, d['e2e']['value'])
P
done
timeout 300 python -m pytest tests/test_npt.py -q -m gpu 2>&1 | tail -3
