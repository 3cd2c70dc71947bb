cd $GRAFT_REPO_ROOT
timeout 800 ncu --set full --clock-control none --import-source on -k regex:tc_fix_a_kernel -s 1 -c 1 -o gpurun_out/r2_prof_fixa python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 8 --csv --log-file gpurun_out/r2_launches_a.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
cut -d, -f5,12- gpurun_out/r2_launches_a.csv | tail -8 | cut -c1-60,120-
