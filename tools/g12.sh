cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "drift or langevin or nlist_r_cut or net_force or tail or pressure_without or num_builds or json" 2>&1 | grep -v "^$" | tail -30
