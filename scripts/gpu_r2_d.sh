mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest-all rc=$?"; tail -n 6 gpurun_out/pytest_gpu_all.log
bash scripts/gpu_r2_multi.sh 2
