cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer1.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_t3.log
cat gpurun_out/r02_t3.log
for v in "" _mb5 _mb4; do
  HB2_LIB=libhoomd_b200$v.so HB2_BENCH_CPU=0 timeout 300 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_quad$v.json 2> gpurun_out/r02_b_quad$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r02_b_quad$v.json').read().strip().splitlines()[-1])
print('$v', d['value'], d['ms_per_step'], d['phases'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'quad_force_kernel|quad_build_kernel' -s 30 -c 4 -o gpurun_out/r02_quad_v1 python bench.py --steps 30 --warmup 5 > gpurun_out/ncu_v1.log 2>&1
tail -3 gpurun_out/ncu_v1.log
