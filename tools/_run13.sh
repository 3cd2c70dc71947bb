cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gmm.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/tc_fullscale_check.py 20000 > gpurun_out/r2b_fs20k.json 2> gpurun_out/r2b_fs20k.err; cat gpurun_out/r2b_fs20k.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])
P
B200_TC_DBG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('alleasy', d['roofline']['kernel_ms'])"
B200_TC_DBG=8 timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nofix', d['roofline']['kernel_ms'])"
