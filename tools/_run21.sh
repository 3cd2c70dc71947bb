cd $GRAFT_REPO_ROOT
echo "== full gpu tests"
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
echo "== smoke"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== compute-sanitizer memcheck"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo rc=$?; tail -3 gpurun_out/r2_sanitizer_memcheck.txt
echo "== compute-sanitizer racecheck"
timeout 900 compute-sanitizer --tool racecheck --print-limit 8 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo rc=$?; tail -3 gpurun_out/r2_sanitizer_racecheck.txt
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 14 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > /dev/null 2>&1
echo "== ncu full score kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_score_kernel -s 6 -c 1 -o gpurun_out/r2_prof_score python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tc_fix -s 4 -c 2 -o gpurun_out/r2_prof_fix python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r2_prof_score.ncu-rep gpurun_out/r2_prof_fix.ncu-rep
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo ref rc=$?
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_whole_step'), d['roofline']['kernel_ms'], d['e2e']['value'])
for k,v in d.get('secondary',{}).items(): print(' ',k, v.get('value'), v.get('ms_per_step'), v.get('roofline',{}).get('frac'))
P
