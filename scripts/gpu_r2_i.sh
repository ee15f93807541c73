# Round-2 pass I: the split-precision chain after the shared-memory-traffic changes (N=64 GEMM2, four H hand-offs).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_bf16_chain.py -m gpu -q -x > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_i.log
timeout 300 python scripts/chain_sweep.py --T 200 --members 1024,4096,18944,37888 --precisions fp32,bf16,bf16x3 --reps 2 2>&1 | tee gpurun_out/chain_sweep_h.log
timeout 300 python scripts/chain_sweep.py --T 1000 --members 1024,18944 --precisions bf16,bf16x3 --reps 1 2>&1 | tee -a gpurun_out/chain_sweep_h.log
timeout 300 python scripts/measure_parity.py > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; grep -A1 "T1000 bf16x3\|cfg1 bf16x3" gpurun_out/parity.log | cut -c1-120
bash scripts/gpu_umma_timing.sh
