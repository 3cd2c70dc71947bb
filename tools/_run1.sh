cd $GRAFT_REPO_ROOT
timeout 600 python tools/tc_fullscale_check.py 8192 2>&1 | tail -1
for dbg in 0 8 9; do
B200_TC_DBG=$dbg timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dbg=$dbg', d['roofline']['kernel_ms'], d['ms_per_step'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/r2_launches_a.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
cut -d, -f5,12- gpurun_out/r2_launches_a.csv | tail -13
