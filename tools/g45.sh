cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | grep real
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','dtype','vs_baseline','gpu_launches')})
print('e2e', d['e2e']); print('roofline', d['roofline']); print('cpu_baseline', d['cpu_baseline']); print('clocks', d['clocks']); print('roofline_step', d['roofline_step'])
P
( time python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_refarm.json 2> gpurun_out/r02_bench_refarm.err ) 2>&1 | grep real
cut -c1-400 gpurun_out/r02_bench_refarm.json
( time timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) 2>&1 | tail -6
