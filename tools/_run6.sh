cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_hmm.py tests/test_z_hmm_big_population.py -x -q -m gpu 2>&1 | tail -4
B200_HMM_PROBE=1 timeout 300 python bench_hmm.py --utts 64 --no-cpu-baseline 2>&1 | grep -E "probe|us_per_frame" | tail -2 | cut -c1-250
B200_HMM_PROBE=1 timeout 300 python bench_hmm.py --utts 1 --frames 400 --warmup 100 --no-cpu-baseline 2>&1 | grep -E "probe|us_per_frame" | tail -2| cut -c1-250
