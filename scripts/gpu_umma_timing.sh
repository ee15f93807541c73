mkdir -p gpurun_out
(cd ert-conditional-diffusion-model_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DUC_TIMING=1 -Xcompiler -fPIC,-fvisibility=hidden -shared -o /tmp/libertdiff_timing.so capi.cu) && \
ERTDIFF_B200_LIB=/tmp/libertdiff_timing.so timeout 300 python scripts/chain_sweep.py --members 18944 --precisions bf16 --timing 2>&1 | tee gpurun_out/sweep_timing.log
