python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep "world=\|differ\|Error\|error" | head -20
