cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_npt.py tests/test_plugin.py -q -m gpu -x -k "lj_2k or c1_64k or npt or pressure or plugin or reference_runs" 2>&1 | tail -3
HB2_BENCH_CPU=0 timeout 200 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_last.json 2> gpurun_out/r02_b_last.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_last.json')); print('last', round(d['ms_per_step'],4), d['e2e']['value'], d['clocks']['samples'])
P
