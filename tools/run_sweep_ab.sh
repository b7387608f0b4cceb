python -m pytest tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -5
ITERS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sweep_launch_v5.csv python tools/time_sweep.py 2>&1 | tail -1
