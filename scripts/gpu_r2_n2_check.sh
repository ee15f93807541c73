# closing 2-GPU sanity pass: parity against a single-GPU run, then the bench line the driver's scaling run launches
mkdir -p gpurun_out
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" | tail -n 12 | tee gpurun_out/multi_gpu_check_$N.log
timeout 900 $RUN --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_n$N.err | tail -n 1 > gpurun_out/bench_n$N.json
tail -n 3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read()); print("bench_n$N", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["config"].get("collectives","")[:50])
for c in d.get("configs", []):
    print("   ", c["name"], "value", round(c["value"]), "ms/step", round(c["ms_per_step"],3), "e2e", round(c["e2e"]["value"]))
PY
