# Round-2 pass H: the split-precision tensor-core chain (precision="bf16x3") and the 4-instruction KDE term.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_chain.py tests/test_gpu_stats.py -m gpu -q > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_h.log
timeout 300 python scripts/measure_parity.py > gpurun_out/parity_h.log 2>&1; echo "parity rc=$?"; grep -i "x3\|umma status" gpurun_out/parity_h.log
timeout 300 python scripts/chain_sweep.py --T 200 --members 1024,4096,18944,37888 --precisions fp32,bf16,bf16x3 --reps 2 2>&1 | tee gpurun_out/chain_sweep_h.log
timeout 300 python scripts/chain_sweep.py --T 1000 --members 1024,18944 --precisions bf16,bf16x3 --reps 1 2>&1 | tee -a gpurun_out/chain_sweep_h.log
timeout 300 python scripts/stats_bench.py --only kde > gpurun_out/stats_bench_kde_h.log 2>&1; echo "stats_bench rc=$?"; cat gpurun_out/stats_bench_kde_h.log
