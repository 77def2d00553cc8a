cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer1.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_t2.log
cat gpurun_out/r02_t2.log
HB2_BENCH_CPU=0 timeout 300 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_quad.json 2> gpurun_out/r02_b_quad.err; tail -c 1800 gpurun_out/r02_b_quad.json; tail -5 gpurun_out/r02_b_quad.err
