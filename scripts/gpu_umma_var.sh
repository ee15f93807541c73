mkdir -p gpurun_out
for V in "2 4" "4 4" "4 8"; do
set -- $V
(cd ert-conditional-diffusion-model_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DUC_TPM_N=$1 -DUC_RNG_WARPS_N=$2 -Xcompiler -fPIC,-fvisibility=hidden -shared -Xptxas -v -o /tmp/libertdiff_v.so capi.cu 2>&1 | grep -A2 "k_chain_ummaILb0ELb0ELb1" | grep "Used\|spill") && \
echo "threads/member $1, RNG warps $2" && ERTDIFF_B200_LIB=/tmp/libertdiff_v.so timeout 300 python scripts/chain_sweep.py --members 18944 --precisions bf16 2>&1 | grep chain_ms && \
ERTDIFF_B200_LIB=/tmp/libertdiff_v.so timeout 300 python scripts/chain_sweep.py --members 18944 --precisions bf16 --distinct 2>&1 | grep chain_ms && \
ERTDIFF_B200_LIB=/tmp/libertdiff_v.so timeout 300 python -m pytest tests/test_gpu_bf16_chain.py -x -q 2>&1 | tail -1
done
