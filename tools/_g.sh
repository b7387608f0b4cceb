cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -2 gpurun_out/bench_r2i.err
python tools/run_k50.py > gpurun_out/k50_r2i.json 2>/dev/null; tail -1 gpurun_out/k50_r2i.json | cut -c90-330
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_sweep_launches_e.csv python tools/time_sweep.py > gpurun_out/ncu_kk_l5.log 2>&1
