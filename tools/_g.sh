cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/run_k50.py 2>&1 | tail -1 | cut -c90-330
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_sweep_launches_d.csv python tools/time_sweep.py > gpurun_out/ncu_kk_l4.log 2>&1
