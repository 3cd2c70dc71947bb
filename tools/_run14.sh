cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_score_kernel -s 6 -c 1 -o gpurun_out/r2d_prof_score python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r2d_prof_score.ncu-rep
