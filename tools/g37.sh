cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_npt.py tests/test_plugin.py -q -m gpu -k "forces_and_energies or family or npt or pressure or plugin or reference_runs" 2>&1 | tail -25
