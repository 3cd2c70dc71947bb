cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_plugin.py -x -q -m gpu -k "sphinx3" 2>&1 | tail -15
