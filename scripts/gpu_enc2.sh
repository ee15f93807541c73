timeout 300 python -m pytest tests/test_gpu_encoder_umma.py tests/test_gpu_bf16_chain.py -x -q 2>&1 | tail -1
timeout 300 python scripts/encoder_bench.py --conds 1,16,64,256,1024,4096 2>&1 | grep bf16
