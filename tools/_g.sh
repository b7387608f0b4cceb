cd $GRAFT_REPO_ROOT
python tools/run_s100k.py 2>&1 | tail -4
