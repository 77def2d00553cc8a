cd $GRAFT_REPO_ROOT
timeout 3000 python -m pytest tests -q -m gpu --durations=12 2>&1 | tail -60 > gpurun_out/r02_t10.log
cat gpurun_out/r02_t10.log
