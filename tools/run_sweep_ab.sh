python -m pytest tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -2
python tools/time_sweep.py 2>&1 | tail -1
python tools/k50_probe.py 2>&1 | tail -6
