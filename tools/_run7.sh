cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_plugin.py tests/test_gpu_gmm.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench_e2e.py --replicas 16 > gpurun_out/r2_e2e_a.json 2> gpurun_out/r2_e2e_a.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_e2e_a.json'))
for m in d['batch']['models']:
    print(m['model'], 'cpu', round(m['cpu']['wall_s'],2), 'plugin', m['plugin'], 'senin', round(m['senin_pipeline']['wall_s'],2))
PY
tail -3 gpurun_out/r2_e2e_a.err
