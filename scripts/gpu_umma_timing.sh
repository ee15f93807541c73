# phase timing of the tensor-core chain from the instrumented build (registers differ: use for proportions only)
mkdir -p gpurun_out
(cd ert-conditional-diffusion-model_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DUC_TIMING=1 -Xcompiler -fPIC,-fvisibility=hidden -shared -o /tmp/libertdiff_timing.so capi.cu) || exit 1
export ERTDIFF_B200_LIB=/tmp/libertdiff_timing.so
for f in "" "--flush"; do
echo "== shared $f"; timeout 120 python scripts/chain_sweep.py --members 8192 --precisions bf16 --timing $f 2>&1
echo "== distinct $f"; timeout 120 python scripts/chain_sweep.py --members 8192 --precisions bf16 --timing --distinct $f 2>&1
done | tee gpurun_out/sweep_timing.log
