cd $GRAFT_REPO_ROOT
CELLS=131072 ITERS=1 python tools/time_sweep.py > gpurun_out/plain_final.log 2>&1 && CELLS=131072 ITERS=1 ncu --set full --clock-control none --import-source on -k regex:"koki_plane|koki_compact" -c 6 -o gpurun_out/prof_final_sweep python tools/time_sweep.py > gpurun_out/ncu_final_sweep.log 2>&1
ls -la gpurun_out/prof_final_sweep.ncu-rep
