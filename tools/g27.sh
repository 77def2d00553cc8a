cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer1.py tests/test_writers.py -q -m gpu -k "not drift" 2>&1 | tail -3
run() { n=$1; shift
  env "$@" timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_$n.json 2> gpurun_out/r02_b_$n.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_$n.json')); print('$n', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases'].items() if isinstance(v,float)})
P
}
run prune1 A=1
run prune0 HB2_TILE_PRUNE=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nlist_build_tile' -s 3 -c 1 -o gpurun_out/r02_tileb_v4 -f python bench.py --steps 30 --warmup 5 > gpurun_out/ncu_tileb_v4.log 2>&1
ncu -i gpurun_out/r02_tileb_v4.ncu-rep --page raw --csv > gpurun_out/r02_tileb_v4_raw.csv; python tools/ncu_summary.py gpurun_out/r02_tileb_v4_raw.csv --all 2>/dev/null | grep -E "=====|time_duration|inst_executed.sum|issue_active|wavefronts.avg|registers|dram__bytes"
