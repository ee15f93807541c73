mkdir -p gpurun_out
python -m pytest tests/test_gpu_stats.py -q 2>&1 | tail -n 2
B="--steps 3 --warmup 3 --no-cpu-baseline --precision bf16 --members 8192"
python bench.py $B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bf16_b8192.csv python bench.py $B > gpurun_out/ncu_b.log 2>&1
