cd $GRAFT_REPO_ROOT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/int_peaks tools/int_peaks.cu && /tmp/int_peaks > gpurun_out/r2_int_peaks.jsonl; cat gpurun_out/r2_int_peaks.jsonl
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json | cut -c1-6000
