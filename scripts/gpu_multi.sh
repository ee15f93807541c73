mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" | tee gpurun_out/multi_gpu_check_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_n$N.err | tail -n 1 > gpurun_out/bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --precision bf16 --members $((8192/N)) 2>gpurun_out/bench_bf16_n$N.err | tail -n 1 > gpurun_out/bench_bf16_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --impl reference --steps 1 --warmup 0 --ref-sample-steps 2 2>/dev/null | tail -n 1 | cut -c1-200
python - <<PY
import json
for f in ["bench_n$N","bench_bf16_n$N"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read()); print(f, "n_gpus", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]))
PY
