python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python tools/run_k50.py --pathways 50 --cells 500000 > gpurun_out/k50_final.json 2> gpurun_out/k50_final.err; cat gpurun_out/k50_final.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
