# Round-2 third GPU pass: tests after the statistics rework, statistics bench, full bench, ncu captures of the statistics kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest-all rc=$?"; tail -n 8 gpurun_out/pytest_gpu_all.log
timeout 600 python scripts/stats_bench.py > gpurun_out/stats_bench.log 2>&1; echo "stats_bench rc=$?"; cat gpurun_out/stats_bench.log
timeout 900 python bench.py > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_r02c.err
S="python scripts/stats_bench.py --maps 1024 --fields 151552 --reps 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_moments -c 1 -f -o gpurun_out/prof_moments $S > gpurun_out/ncu_s1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_percentiles_warp -c 1 -f -o gpurun_out/prof_percentiles $S > gpurun_out/ncu_s2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_kde_scan32 -c 1 -f -o gpurun_out/prof_kde_scan $S > gpurun_out/ncu_s3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_kde_select64 -c 1 -f -o gpurun_out/prof_kde_select $S > gpurun_out/ncu_s4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_posterior_update -c 1 -f -o gpurun_out/prof_posterior_update $S > gpurun_out/ncu_s5.log 2>&1
ls -la gpurun_out/*.ncu-rep
