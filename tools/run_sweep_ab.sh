timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sweep.py -x -q -m gpu -k "golden or long_cycles or (general_tables and 65-1000) or trajectory_distances" 2>&1 | tail -6
echo "memcheck rc=$?"
SCB_SWEEP_FAST_STEPS=5 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_sweep.py -x -q -m gpu -k "sweep_golden or (random_vs_oracle and 17-100)" 2>&1 | tail -6
