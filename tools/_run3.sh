cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
python tools/time_sweep.py 2>&1 | tail -1
SCB_SWEEP_SORT=0 python tools/time_sweep.py 2>&1 | tail -1
