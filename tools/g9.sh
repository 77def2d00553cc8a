cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -q -m gpu -x --durations=8 2>&1 | tail -40 > gpurun_out/r02_t9.log
cat gpurun_out/r02_t9.log
timeout 900 python bench.py --steps 1000 --warmup 100 > gpurun_out/r02_bench_1gpu_a.json 2> gpurun_out/r02_bench_1gpu_a.err
tail -c 2500 gpurun_out/r02_bench_1gpu_a.json; tail -3 gpurun_out/r02_bench_1gpu_a.err
