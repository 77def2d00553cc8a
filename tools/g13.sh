cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer1.py -x -q -m gpu -k "not drift" 2>&1 | tail -40 > gpurun_out/r02_t13.log
cat gpurun_out/r02_t13.log
HB2_BENCH_CPU=0 timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_tile1.json 2> gpurun_out/r02_b_tile1.err; tail -c 2500 gpurun_out/r02_b_tile1.json; tail -5 gpurun_out/r02_b_tile1.err
