cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_s3_hmm.py tests/test_gpu_hmm.py tests/test_plugin.py -x -q -m gpu 2>&1 | tail -12
