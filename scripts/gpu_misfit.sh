mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_misfit_metrics.py -x -q 2>&1 | tail -15
timeout 300 python scripts/misfit_bench.py 2>&1 | tee gpurun_out/misfit_bench.log
