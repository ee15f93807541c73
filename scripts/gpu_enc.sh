mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_encoder_umma.py -x -q 2>&1 | tail -3
timeout 300 python scripts/encoder_bench.py 2>&1 | tee gpurun_out/encoder_bench.log
ENC_TIMING=1 ERTDIFF_B200_LIB=/root/repo/scripts/_dbg/lib_t.so timeout 300 python scripts/encoder_bench.py --conds 1,1024 2>&1 | grep cycles
