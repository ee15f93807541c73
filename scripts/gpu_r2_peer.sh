# peer-memory all-gather: parity against NCCL, then the default bench with both collectives, then the breakdown
mkdir -p gpurun_out
N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" | tail -n 12 | tee gpurun_out/multi_gpu_check_$N.log
timeout 300 $RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-extra-configs 2>gpurun_out/bench_peer_n$N.err | tail -n 1 > gpurun_out/bench_peer_n$N.json
ERTDIFF_BENCH_NCCL=1 timeout 300 $RUN --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-extra-configs 2>gpurun_out/bench_nccl_n$N.err | tail -n 1 > gpurun_out/bench_nccl_n$N.json
timeout 300 $RUN --master-port 29520 scripts/multi_gpu_breakdown.py 2>gpurun_out/breakdown_n$N.err | grep "^#\|^|" > gpurun_out/breakdown_fp32_256_n$N.md
tail -n 4 gpurun_out/bench_peer_n$N.err
python - <<PY
import json
for f in ["bench_peer_n$N","bench_nccl_n$N"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read()); print(f, "n_gpus", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["config"].get("collectives","")[:40])
    except Exception as e: print(f, "failed", e)
PY
head -16 gpurun_out/breakdown_fp32_256_n$N.md | cut -c1-140
