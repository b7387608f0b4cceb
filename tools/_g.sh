cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_rules.py tests/test_gpu_entry_points.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for pw in 3 4 5; do echo -n "pw=$pw"; SCB_FITNESS_PER_WARP=$pw python bench.py --steps 40 --warmup 5 --sweep-steps 0 --no-cpu-baseline --no-e2e --no-e2e-dense 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   step %.2f us  counts alone %.2f us  checksum %d' % (d['ms_per_step']*1e3, d['roofline']['kernel_ms']*1e3, d['config']['fitness_checksum']))"; done
