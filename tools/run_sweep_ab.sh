python -m pytest tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -2
python tools/time_sweep.py 2>&1 | tail -1
SCB_SWEEP_THREADS=1024 python -m scbonita2_b200.build --force | tail -1
python -m pytest tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -2
SCB_SWEEP_FAST_STEPS=6 python -m pytest tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -2
python tools/time_sweep.py 2>&1 | tail -1
SCB_TRACE=1 SCB_SWEEP_THREADS=1024 python -m scbonita2_b200.build --force | tail -1
python tools/trace_sweep.py 2>&1 | tail -8 | head -4
