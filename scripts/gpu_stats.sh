python -m pytest tests/test_gpu_stats.py -q 2>&1 | tail -n 8
