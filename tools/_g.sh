cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -2 gpurun_out/bench_r2f.err
CELLS=131072 ITERS=1 ncu --set full --clock-control none --import-source on -k regex:koki_compact -c 2 -o gpurun_out/prof_kc python tools/time_sweep.py > gpurun_out/ncu_kc.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_sweep_launches_c.csv python tools/time_sweep.py > gpurun_out/ncu_kk_l3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k50_launches_c.csv python tools/run_k50.py > gpurun_out/ncu_k50_l3.log 2>&1
