python -m pytest tests/test_gpu_stats.py tests/test_uq_calibration.py -q 2>&1 | tail -n 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python scripts/step_breakdown.py
bash scripts/gpu_quick_bench.sh | grep bench
