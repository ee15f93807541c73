# Round-2 last single-GPU pass (after the coarse-to-fine KDE scan): the whole test suite, smoke, the bench line, the statistics bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest-all rc=$?"; tail -n 2 gpurun_out/pytest_gpu_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"; tail -n 2 gpurun_out/bench_r02_final.err
timeout 200 python scripts/stats_bench.py > gpurun_out/stats_bench.log 2>&1; echo "stats_bench rc=$?"; grep KDE gpurun_out/stats_bench.log
