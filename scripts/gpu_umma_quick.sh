mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16_chain.py tests/test_gpu_denoiser.py -x -q 2>&1 | tail -3
timeout 600 python scripts/chain_sweep.py --members 256,1024,18944,37888 2>&1 | tee gpurun_out/sweep_quick.log
timeout 300 python scripts/chain_sweep.py --distinct --members 18944 --precisions bf16 2>&1 | tee -a gpurun_out/sweep_quick.log
# phase timing from the instrumented build (registers differ: use for proportions only)
(cd ert-conditional-diffusion-model_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DUC_TIMING=1 -Xcompiler -fPIC,-fvisibility=hidden -shared -o /tmp/libertdiff_timing.so capi.cu) && \
ERTDIFF_B200_LIB=/tmp/libertdiff_timing.so timeout 300 python scripts/chain_sweep.py --members 18944 --precisions bf16 --timing 2>&1 | tee -a gpurun_out/sweep_quick.log
