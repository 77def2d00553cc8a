cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_bonds.py tests/test_gpu_layer1.py -q -m gpu 2>&1 | tail -4
for o in rowmajor morton; do
  HB2_CELL_ORDER=$o HB2_BENCH_CPU=0 timeout 300 python bench.py --steps 600 --warmup 100 > gpurun_out/r02_b_order_$o.json 2> gpurun_out/r02_b_order_$o.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r02_b_order_$o.json').read().strip().splitlines()[-1])
print('$o', d['value'], d['ms_per_step'], d['phases'], d['parity']['pass'])
PY
  HB2_CELL_ORDER=$o HB2_BENCH_CPU=0 timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none -k regex:'pair_force_kernel|nlist_build_kernel' -s 30 -c 8 --csv --log-file gpurun_out/r02_order_$o.csv python bench.py --steps 30 --warmup 5 > /dev/null 2>&1
done
