# Round-2 GPU pass e: tests after the statistics changes, statistics bench, default bench, ncu of the new kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest-all rc=$?"; tail -n 8 gpurun_out/pytest_gpu_all.log
timeout 600 python scripts/stats_bench.py > gpurun_out/stats_bench.log 2>&1; echo "stats_bench rc=$?"; cat gpurun_out/stats_bench.log
timeout 900 python bench.py > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_r02e.err
S="python scripts/stats_bench.py --maps 1024 --fields '' --reps 0 --only pct"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_percentiles_select -c 1 -f -o gpurun_out/prof_percentiles python scripts/stats_bench.py --maps 1024 --fields "" --reps 0 --only pct > gpurun_out/ncu_s2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_percentiles_warp -c 1 -f -o gpurun_out/prof_percentiles_n50 python scripts/stats_bench.py --maps 50 --fields "" --reps 0 --only pct > gpurun_out/ncu_s6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sort_runs -c 1 -f -o gpurun_out/prof_sort_runs python scripts/stats_bench.py --maps "" --fields 151552 --reps 0 --only pct > gpurun_out/ncu_s7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_select_runs -c 1 -f -o gpurun_out/prof_select_runs python scripts/stats_bench.py --maps "" --fields 151552 --reps 0 --only pct > gpurun_out/ncu_s8.log 2>&1
D="python scripts/chain_sweep.py --members 18944 --precisions bf16 --T 200 --reps 1"
timeout 300 $D > gpurun_out/plain_d.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_umma -s 1 -c 1 -f -o gpurun_out/prof_chain_umma $D > gpurun_out/ncu_d.log 2>&1
E="python scripts/encoder_bench.py --conds 1024 --reps 1"
timeout 300 $E > gpurun_out/plain_e.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_encoder_umma -s 1 -c 1 -f -o gpurun_out/prof_encoder_umma $E > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep
