cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer1.py tests/test_bonds.py tests/test_writers.py tests/test_gsd_io.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r02_t6.log
cat gpurun_out/r02_t6.log
