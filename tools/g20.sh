cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer1.py tests/test_writers.py -x -q -m gpu -k "not drift" 2>&1 | tail -5
run() { # name, env...
  n=$1; shift
  env "$@" timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_$n.json 2> gpurun_out/r02_b_$n.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_$n.json')); print('$n', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases'].items() if isinstance(v,float)})
P
}
run base A=1
run nofuse HB2_FUSE_INTEGRATE=0
run minb4 HB2_TILE_BUILD_MINB=4
run pf0 HB2_TILE_PREFETCH=0
run pf1 HB2_TILE_PREFETCH=1
run pf2 HB2_TILE_PREFETCH=2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_force_tile' -s 30 -c 4 -o gpurun_out/r02_tilef_v10 -f python bench.py --steps 30 --warmup 5 > gpurun_out/ncu_tilef_v10.log 2>&1
ncu -i gpurun_out/r02_tilef_v10.ncu-rep --page raw --csv > gpurun_out/r02_tilef_v10_raw.csv; python tools/ncu_summary.py gpurun_out/r02_tilef_v10_raw.csv --all 2>/dev/null | grep -E "=====|time_duration|inst_executed.sum|issue_active|wavefronts.avg|long_score|dram__bytes"
