mkdir -p gpurun_out
python -m pytest tests/test_gpu_stats.py -q 2>&1 | tail -n 3
python bench.py --steps 10 --warmup 3 --precision bf16 --members 8192 --no-cpu-baseline > gpurun_out/bench_bf16_8192.json 2> gpurun_out/bench_bf16_8192.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32_nocpu.json 2>gpurun_out/bench_fp32_nocpu.err
python - <<'PY'
import json
for f in ["bench_bf16_8192","bench_fp32_nocpu"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"])
PY
