cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer1.py -x -q -m gpu -k "not drift" 2>&1 | tail -15
timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_tile8.json 2> gpurun_out/r02_b_tile8.err; python - <<'P'
import json; d=json.load(open('gpurun_out/r02_b_tile8.json')); print(d['ms_per_step'], d['phases'])
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_force_tile' -s 20 -c 1 -o gpurun_out/r02_tilef_v8 -f python bench.py --steps 30 --warmup 5 > gpurun_out/ncu_tilef_v8.log 2>&1
for r in r02_tilef_v8; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv; python tools/ncu_summary.py gpurun_out/${r}_raw.csv --all 2>/dev/null | grep -E "=====|time_duration|inst_executed.sum|issue_active|registers|wavefronts.avg|warps_active|short_score|long_score|barrier|not_selected|wait"; done
