# final multi-GPU pass (N = $1): parity, peer-vs-NCCL on the default workload, per-kernel breakdown, the full bench line
mkdir -p gpurun_out
N=${1:-8}
bash scripts/gpu_r2_peer.sh $N
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_n$N.err | tail -n 1 > gpurun_out/bench_n$N.json
timeout 300 $RUN --master-port 29521 scripts/multi_gpu_breakdown.py --members 1024 --precision bf16 2>>gpurun_out/breakdown_n$N.err | grep "^#\|^|" > gpurun_out/breakdown_bf16_1024_n$N.md
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read()); print("bench_n$N", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]))
for c in d.get("configs", []):
    print("   ", c["name"], "value", round(c["value"]), "ms/step", round(c["ms_per_step"],3), "e2e", round(c["e2e"]["value"]))
PY
