cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-dense > gpurun_out/plain_final2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-dense > gpurun_out/ncu_final_l2.log 2>&1
tail -c 300 gpurun_out/plain_final2.log
