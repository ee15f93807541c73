mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_chain.py tests/test_gpu_stats.py -m gpu -q -x > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_f.log
timeout 300 python scripts/chain_sweep.py --members 1024,4096,8192,18944,37888 --precisions bf16 > gpurun_out/sweep_bf16.log 2>&1; cat gpurun_out/sweep_bf16.log
timeout 300 python scripts/chain_sweep.py --members 8192,18944 --precisions bf16 --distinct > gpurun_out/sweep_bf16_distinct.log 2>&1; cat gpurun_out/sweep_bf16_distinct.log
timeout 600 python scripts/stats_bench.py --maps 1024,2048,8192 --fields "" --only pct > gpurun_out/stats_pct_select.log 2>&1; cat gpurun_out/stats_pct_select.log
ERTDIFF_PCTL_NO_SELECT=1 timeout 600 python scripts/stats_bench.py --maps 1024,2048,8192 --fields "" --only pct > gpurun_out/stats_pct_noselect.log 2>&1; cat gpurun_out/stats_pct_noselect.log
ERTDIFF_PCTL_SELECT=1 timeout 600 python scripts/stats_bench.py --maps 512,1024 --fields "" --only pct > gpurun_out/stats_pct_forced.log 2>&1; cat gpurun_out/stats_pct_forced.log
