cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_hmm.py tests/test_z_hmm_big_population.py -x -q -m gpu 2>&1 | tail -3
B200_HMM_PROBE=1 timeout 300 python bench_hmm.py --utts 64 --frames 200 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-200
timeout 300 python bench_hmm.py --utts 64 --frames 200 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batched', d['us_per_frame'], d['roofline']['frac'], d['survivor_fraction'])"
timeout 300 python bench_hmm.py --utts 1 --frames 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('single', d['us_per_frame'], d['roofline']['frac'], d['survivor_fraction'])"
