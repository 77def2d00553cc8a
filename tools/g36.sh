cd $GRAFT_REPO_ROOT
timeout 600 python tests/plugin_check.py 2>&1 | tail -15
