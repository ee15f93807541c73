mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest-all rc=$?"; tail -n 5 gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench rc=$?"
python bench.py --steps 10 --warmup 3 --distinct-conditions --no-cpu-baseline > gpurun_out/bench_fp32_distinct.json 2> gpurun_out/bench_fp32_distinct.err
python bench.py --steps 10 --warmup 3 --precision bf16 --members 1024 --no-cpu-baseline > gpurun_out/bench_bf16_1024.json 2> gpurun_out/bench_bf16_1024.err
python bench.py --steps 10 --warmup 3 --precision bf16 --members 8192 --no-cpu-baseline > gpurun_out/bench_bf16_8192.json 2> gpurun_out/bench_bf16_8192.err
python bench.py --steps 10 --warmup 3 --precision bf16 --members 8192 --distinct-conditions --no-cpu-baseline > gpurun_out/bench_bf16_8192_distinct.json 2> gpurun_out/bench_bf16_8192_distinct.err
python bench.py --steps 10 --warmup 3 --members 8192 --no-cpu-baseline > gpurun_out/bench_fp32_8192.json 2> gpurun_out/bench_fp32_8192.err
tail -n 2 gpurun_out/*.err
