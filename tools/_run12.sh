cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 12 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_score_kernel -s 6 -c 1 -o gpurun_out/r2_prof_score python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:hmm_run_kernel -c 1 -o gpurun_out/r2_prof_hmm python bench_hmm.py --utts 64 --frames 40 --warmup 8 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo ref rc=$?; cut -c1-300 gpurun_out/r2_bench_reference.json
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
