python -m pytest tests/test_gpu_stats.py tests/test_uq_calibration.py -q 2>&1 | tail -n 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
for m in 256 1024 2048; do python scripts/step_breakdown.py --members $m | grep kde; done
python scripts/step_breakdown.py --members 8192 --precision bf16 | grep "kde\|moments\|percentile"
