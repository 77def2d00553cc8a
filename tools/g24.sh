cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 2700 python -m pytest tests -q -m gpu 2>&1 | tail -12
