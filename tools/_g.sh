cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; tail -2 gpurun_out/bench_r2e.err
CELLS=131072 ITERS=1 ncu --set full --clock-control none --import-source on -k regex:koki_plane -c 1 -o gpurun_out/prof_kk_d python tools/time_sweep.py > gpurun_out/ncu_kk_d.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_sweep_launches_b.csv python tools/time_sweep.py > gpurun_out/ncu_kk_l2.log 2>&1
