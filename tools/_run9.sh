cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload e2e_decode --gpus 2 --utts-per-gpu 128 --cpu-sample 64 2> gpurun_out/cfg5.err | tee gpurun_out/r2_cfg5_2gpu.json | cut -c1-400
tail -3 gpurun_out/cfg5.err
