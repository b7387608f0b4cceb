python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -c 600 gpurun_out/bench8.err; cat gpurun_out/bench8.json
