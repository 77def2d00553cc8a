cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bonds.py -q -m gpu -x -k "not drift and not forces_and_energies and not family" 2>&1 | tail -5
for ch in 0 8 4 12 6; do
HB2_TILE_CHUNK=$ch timeout 300 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_chunk$ch.json 2> gpurun_out/r02_b_chunk$ch.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_chunk$ch.json')); print('chunk=$ch', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases'].items() if isinstance(v,float)}, d['e2e']['value'])
P
done
