set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -15
for so in 24 8 64; do for ss in 4 2; do
SCB_SWEEP_STOP_OPEN=$so SCB_SWEEP_STOP_STALL=$ss python tools/time_sweep.py 2>&1 | tail -1 | sed "s/^/open=$so stall=$ss: /"
done; done
SCB_SWEEP_SORT=0 python tools/time_sweep.py 2>&1 | tail -1 | sed "s/^/unsorted: /"
