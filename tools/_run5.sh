cd $GRAFT_REPO_ROOT
python tools/run_k50.py --pathways 50 --cells 500000 2>&1 | tail -1
SCB_SWEEP_SORT=0 python tools/run_k50.py --pathways 50 --cells 500000 2>&1 | tail -1
