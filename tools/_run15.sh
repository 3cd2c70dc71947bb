cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fix -s 4 -c 2 -o gpurun_out/r2e_prof_fix python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r2e_prof_fix.ncu-rep
timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])
P
B200_TC_DBG=8 timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nofix', d['roofline']['kernel_ms'])"
