cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
ITERS=3 python tools/time_sweep.py | tail -1 | cut -c1-40
CELLS=125000 ITERS=3 python tools/time_sweep.py | tail -1 | cut -c1-40
