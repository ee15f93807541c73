# phase timing of the tensor-core chain from the instrumented build (registers differ: use for proportions only).
# Build it first, here (it travels with the snapshot):
#   cd ert-conditional-diffusion-model_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DUC_TIMING=1 \
#     -Xcompiler -fPIC,-fvisibility=hidden -c -o /tmp/chain_umma_timing.o chain_umma.cu && nvcc -gencode arch=compute_100a,code=sm_100a \
#     -shared -o ../libertdiff_timing.so _obj/capi.o _obj/chain_fp32.o /tmp/chain_umma_timing.o _obj/encoder.o _obj/stats.o _obj/peer.o
mkdir -p gpurun_out
export ERTDIFF_B200_LIB=$PWD/ert-conditional-diffusion-model_b200/libertdiff_timing.so
timeout 200 python scripts/chain_sweep.py --members 18944,4096 --precisions bf16,bf16x3 --T 200 --reps 2 --timing 2>&1 | tee gpurun_out/sweep_timing.log
