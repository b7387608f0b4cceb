cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py tests/test_gpu_entry_points.py -x -q -m gpu 2>&1 | tail -3
python tools/time_sweep.py | tail -1
