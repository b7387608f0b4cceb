cd $GRAFT_REPO_ROOT
python tools/time_sweep.py 2>&1 | tail -1 | sed 's/^/arity-aware: /'
SCB_DEFINES="SCB_RING_ARITY=0" python -m scbonita2_b200.build --force > /dev/null
python tools/time_sweep.py 2>&1 | tail -1 | sed 's/^/branch-free: /'
SCB_TRACE=1 python -m scbonita2_b200.build --force > /dev/null
python tools/trace_sweep.py 2>&1 | tail -12
