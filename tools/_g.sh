cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
SCB_SWEEP_RECHECK=old python tools/time_sweep.py | tail -1
python tools/time_sweep.py | tail -1
python tools/run_k50.py 2>&1 | tail -1
