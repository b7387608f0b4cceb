cd $GRAFT_REPO_ROOT
cp scbonita2_b200/csrc/libscbonita_b200.so /tmp/keep.so
cp tools/exp/_libs/lib_trace.so scbonita2_b200/csrc/libscbonita_b200.so
python tools/trace_pdl.py | tail -22
cp /tmp/keep.so scbonita2_b200/csrc/libscbonita_b200.so
