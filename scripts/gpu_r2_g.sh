mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_all.log
timeout 300 python scripts/summary_window_bench.py 2>/dev/null | tee gpurun_out/summary_window.log
timeout 600 python scripts/stats_bench.py --maps 50,1024,8192 --fields "" --only pct 2>&1 | tee gpurun_out/stats_pct.log
timeout 900 python bench.py > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_r02g.err
