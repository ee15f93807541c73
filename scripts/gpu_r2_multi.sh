# Round-2 multi-GPU pass (N = $1): parity of the sharded run, the bench with its sub-configs, the per-kernel breakdown.
mkdir -p gpurun_out
N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" | tee gpurun_out/multi_gpu_check_$N.log
timeout 900 $RUN --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_n$N.err | tail -n 1 > gpurun_out/bench_n$N.json
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --precision bf16 --members 18944 --no-extra-configs 2>gpurun_out/bench_bf16_18944_n$N.err | tail -n 1 > gpurun_out/bench_bf16_18944_n$N.json
timeout 300 $RUN --master-port 29520 scripts/multi_gpu_breakdown.py 2>gpurun_out/breakdown_n$N.err | grep "^#\|^|" > gpurun_out/breakdown_fp32_256_n$N.md
timeout 300 $RUN --master-port 29521 scripts/multi_gpu_breakdown.py --members 1024 --precision bf16 2>>gpurun_out/breakdown_n$N.err | grep "^#\|^|" > gpurun_out/breakdown_bf16_1024_n$N.md
tail -n 3 gpurun_out/*_n$N.err
python - <<PY
import json
for f in ["bench_n$N","bench_bf16_18944_n$N"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read()); print(f, "n_gpus", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "chain", d.get("roofline",{}).get("kernel_ms"))
    for c in d.get("configs", []):
        print("   ", c["name"], "value", round(c["value"]), "ms/step", round(c["ms_per_step"],3), "e2e", round(c["e2e"]["value"]))
PY
head -30 gpurun_out/breakdown_fp32_256_n$N.md
