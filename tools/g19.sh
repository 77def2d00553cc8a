cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
V=${V:-9}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer1.py -x -q -m gpu -k "not drift" 2>&1 | tail -5
timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_tile$V.json 2> gpurun_out/r02_b_tile$V.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_tile$V.json')); print(d['ms_per_step'], d['phases'])
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_force_tile|nlist_build_tile' -s 40 -c 2 -o gpurun_out/r02_tile_v$V -f python bench.py --steps 30 --warmup 5 > gpurun_out/ncu_tile_v$V.log 2>&1
ncu -i gpurun_out/r02_tile_v$V.ncu-rep --page raw --csv > gpurun_out/r02_tile_v${V}_raw.csv; python tools/ncu_summary.py gpurun_out/r02_tile_v${V}_raw.csv --all 2>/dev/null | grep -E "=====|time_duration|inst_executed.sum|issue_active|registers|wavefronts.avg|warps_active|short_score|long_score|barrier|not_selected|wait|dram__bytes"
