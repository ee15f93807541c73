python -m pytest tests/test_gpu_denoiser.py tests/test_gpu_encoder_umma.py -q 2>&1 | tail -n 2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32_nocpu.json 2>gpurun_out/bench_fp32_nocpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_fp32_nocpu.json").read().strip().splitlines()[-1]); print("bench", d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"])
PY
