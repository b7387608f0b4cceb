cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -4
SCB_SWEEP_SINGLES=0 ITERS=2 python tools/time_sweep.py | tail -1
ITERS=2 python tools/time_sweep.py | tail -1
