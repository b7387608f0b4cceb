cd $GRAFT_REPO_ROOT
for m in 0 1 0 1; do echo -n "mid=$m "; SCB_SWEEP_BRENT_MID=$m python tools/run_k50.py 2>&1 | tail -1 | cut -c230-290; done
for m in 0 1; do echo -n "mid=$m "; SCB_SWEEP_BRENT_MID=$m ITERS=3 python tools/time_sweep.py | tail -1 | cut -c1-40; done
