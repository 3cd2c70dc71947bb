cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== compute-sanitizer memcheck"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo rc=$?; tail -4 gpurun_out/r2_sanitizer_memcheck.txt
echo "== compute-sanitizer racecheck"
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo rc=$?; tail -4 gpurun_out/r2_sanitizer_racecheck.txt
