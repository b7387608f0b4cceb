cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -2 gpurun_out/bench_r2d.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-dense > gpurun_out/plain_r2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-dense > gpurun_out/ncu_r2_launch.log 2>&1
python tools/profile_kernels.py --what counts,fitness --reps 3 > gpurun_out/plain_r2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rule_counts_kernel -s 2 -c 1 -o gpurun_out/prof_counts_r2_h1m python tools/profile_kernels.py --what counts,fitness --reps 3 > gpurun_out/ncu_r2_counts.log 2>&1
CELLS=131072 ITERS=1 python tools/time_sweep.py > gpurun_out/plain_r2c.log 2>&1 &&
CELLS=131072 ITERS=1 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -c 6 -o gpurun_out/prof_sweep_r2 python tools/time_sweep.py > gpurun_out/ncu_r2_sweep.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
