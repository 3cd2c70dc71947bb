cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_s3.py tests/test_s3_hmm.py -x -q -m gpu 2>&1 | tail -15
