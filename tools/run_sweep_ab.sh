python bench.py --steps 2 --warmup 3 > gpurun_out/b_pre_ncu2.json 2> gpurun_out/b_pre_ncu2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v6.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v6.log 2>&1
tail -2 gpurun_out/ncu_v6.log | cut -c1-300
