python bench_hmm.py --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batched us/frame', d['us_per_frame'], 'frac', d['roofline']['frac'], d['survivor_fraction'])"
python bench_hmm.py --utts 1 --frames 2000 --warmup 200 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('single us/frame', d['us_per_frame'])"
