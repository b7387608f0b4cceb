cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2i_n8.json 2> gpurun_out/bench_r2i_n8.err; tail -1 gpurun_out/bench_r2i_n8.err | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/run_k50.py > gpurun_out/k50_r2i_n8.json 2> gpurun_out/k50_r2i_n8.err; tail -1 gpurun_out/k50_r2i_n8.json
