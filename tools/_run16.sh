cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_hmm.py tests/test_z_hmm_big_population.py tests/test_s3_hmm.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench_hmm.py --utts 64 --frames 200 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batched', d['us_per_frame'], d['roofline']['frac'], d['survivor_fraction'])"
timeout 300 python bench_hmm.py --utts 1 --frames 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('single', d['us_per_frame'], d['roofline']['frac'], d['survivor_fraction'])"
B200_HMM_CLUSTER=0 timeout 300 python bench_hmm.py --utts 64 --frames 200 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batched coop', d['us_per_frame'], d['roofline']['frac'])"
B200_HMM_CLUSTER=0 timeout 300 python bench_hmm.py --utts 1 --frames 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('single coop', d['us_per_frame'], d['roofline']['frac'])"
