cd $GRAFT_REPO_ROOT
SCB_SWEEP_KK_CUR=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/kk_launches.csv python tools/time_sweep.py > gpurun_out/ncu_kk_l.log 2>&1
CELLS=131072 ITERS=1 SCB_SWEEP_KK_CUR=0 ncu --set full --clock-control none --import-source on -k regex:koki_plane -c 1 -o gpurun_out/prof_kk_c python tools/time_sweep.py > gpurun_out/ncu_kk_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -1
