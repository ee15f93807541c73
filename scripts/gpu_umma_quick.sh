mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_chain.py tests/test_gpu_denoiser.py -x -q 2>&1 | tail -15
for m in 32 64; do ERTDIFF_UMMA_MPC=$m timeout 600 python -m pytest tests/test_gpu_bf16_chain.py -x -q 2>&1 | tail -3; done
timeout 600 python scripts/chain_sweep.py --precisions bf16 --members 256,1024,4096,8192,9472,12288 2>&1 | tee gpurun_out/sweep_quick.log
timeout 300 python scripts/chain_sweep.py --distinct --members 1024,8192 --precisions bf16 2>&1 | tee -a gpurun_out/sweep_quick.log
