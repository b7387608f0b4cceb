cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -2 gpurun_out/bench_r2h.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2h_ref.json 2> gpurun_out/bench_r2h_ref.err; tail -c 600 gpurun_out/bench_r2h_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-dense > gpurun_out/ncu_final_l.log 2>&1
