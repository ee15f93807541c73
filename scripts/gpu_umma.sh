mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_chain.py tests/test_gpu_umma.py tests/test_gpu_denoiser.py -x -q > gpurun_out/pytest_umma.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_umma.log
timeout 600 python scripts/chain_sweep.py 2>&1 | tee gpurun_out/sweep.log
timeout 300 python scripts/chain_sweep.py --distinct --members 8192,37888 2>&1 | tee gpurun_out/sweep_distinct.log
