mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_bf16_chain.py -x -q --timeout 30 2>&1 | tail -3
for lib in scripts/_dbg/lib_nopf.so ert-conditional-diffusion-model_b200/libertdiff_b200.so; do
echo "== $lib" | tee -a gpurun_out/sweep_quick.log
for f in "" "--flush"; do
ERTDIFF_B200_LIB=$PWD/$lib timeout 90 python scripts/chain_sweep.py $f --precisions bf16 --members 1024,8192 2>&1 | tee -a gpurun_out/sweep_quick.log
done
done
