mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_bf16_chain.py -x -q --timeout 30 2>&1 | tail -3
timeout 90 python scripts/chain_sweep.py --precisions bf16 --members 256,1024,4096,8192,9472,18944,37888 2>&1 | tee gpurun_out/sweep_quick.log
timeout 90 python scripts/chain_sweep.py --distinct --members 8192,37888 --precisions bf16 2>&1 | tee -a gpurun_out/sweep_quick.log
