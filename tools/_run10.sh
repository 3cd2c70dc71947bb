cd $GRAFT_REPO_ROOT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload e2e_decode --gpus 8 2> gpurun_out/cfg5_8.err | tee gpurun_out/r2_cfg5_8gpu.json | cut -c1-300
tail -3 gpurun_out/cfg5_8.err
