cd $GRAFT_REPO_ROOT
which nvidia-cuda-mps-control nvidia-cuda-mps-server 2>&1
nproc
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
timeout 30 nvidia-cuda-mps-control -d; echo "mps start rc=$?"
sleep 1
timeout 900 python bench_e2e.py --replicas 16 > gpurun_out/r2_e2e_mps.json 2> gpurun_out/r2_e2e_mps.err; echo "bench rc=$?"
echo quit | timeout 30 nvidia-cuda-mps-control; echo "mps quit rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_e2e_mps.json').read().strip().splitlines()[-1])
for m in d['batch']['models']:
    print(m['model'], 'cpu', round(m['cpu']['wall_s'],2), 'plugin16', round(m['plugin']['wall_s'],2), 'plugin4', round(m['plugin']['with_4_processes']['wall_s'],2), 'senin', round(m['senin_pipeline']['wall_s'],2), m['plugin']['identical_words'], m['plugin'].get('one_process_one_utterance_wall_s'))
P
tail -5 /tmp/mps_log/control.log 2>/dev/null
