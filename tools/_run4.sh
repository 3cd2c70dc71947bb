cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
timeout 900 python tools/tc_fullscale_check.py 100000 > gpurun_out/r2_tc_fullscale_parity.json 2>gpurun_out/fs.err; cat gpurun_out/r2_tc_fullscale_parity.json
B200_TC_F16=0 timeout 900 python tools/tc_fullscale_check.py 100000 > gpurun_out/r2_tc_fullscale_parity_tf32.json 2>>gpurun_out/fs.err; cat gpurun_out/r2_tc_fullscale_parity_tf32.json
