cd $GRAFT_REPO_ROOT
HB2_FUSE_INTEGRATE=0 timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "drift_1e5_reference" -s 2>&1 | grep -E "NVE|passed|failed"
HB2_TILE=0 timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "drift_1e5_reference" -s 2>&1 | grep -E "NVE|passed|failed"
