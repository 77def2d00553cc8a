cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "lj_2k or c1_64k or true_pair or edge or ortho or ka_2k" 2>&1 | tail -12
for loc in 1 0; do
HB2_TILE_LOCAL=$loc timeout 200 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_local$loc.json 2> gpurun_out/r02_b_local$loc.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_local$loc.json')); print('local=$loc', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases'].items() if isinstance(v,float)}, d['e2e']['value'])
P
tail -2 gpurun_out/r02_b_local$loc.err
done
