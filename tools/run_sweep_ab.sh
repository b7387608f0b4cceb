python -m pytest tests/test_gpu_rules.py -x -q -m gpu 2>&1 | tail -3
