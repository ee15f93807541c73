# Round-2 first GPU pass: tests, parity measurement, kernel-variant sweeps, bench.  Every step under `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_stats.py::test_statistics_full_size_151k_members > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_gpu_a.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest-all rc=$?"; tail -n 25 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 600 python scripts/measure_parity.py > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -n 60 gpurun_out/parity.log
timeout 300 python scripts/chain_fp32_variants.py > gpurun_out/variants_h128.log 2>&1; echo "variants rc=$?"; cat gpurun_out/variants_h128.log
timeout 300 python scripts/chain_fp32_variants.py --hidden 256 --members 256,512,4096 --L 9386 > gpurun_out/variants_h256.log 2>&1; cat gpurun_out/variants_h256.log
timeout 300 python scripts/chain_sweep.py --members 1024,8192,18944,37888 --precisions bf16 > gpurun_out/sweep_bf16.log 2>&1; cat gpurun_out/sweep_bf16.log
timeout 900 python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_r02a.err; head -c 1500 gpurun_out/bench_r02a.json
