python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
SCB_SWEEP_FAST_STEPS=6 python -m pytest tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -2
python tools/time_sweep.py 2>&1 | tail -1
python tools/k50_probe.py 2>&1 | tail -6
