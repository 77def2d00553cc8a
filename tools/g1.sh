set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer1.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_t1.log
cat gpurun_out/r02_t1.log
HB2_BENCH_CPU=0 timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_quad.json 2> gpurun_out/r02_b_quad.err; tail -c 3000 gpurun_out/r02_b_quad.json; tail -5 gpurun_out/r02_b_quad.err
HB2_LIST=rows HB2_BENCH_CPU=0 timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_rows.json 2> gpurun_out/r02_b_rows.err; tail -c 1500 gpurun_out/r02_b_rows.json
