mkdir -p gpurun_out
CMD="python scripts/encoder_bench.py --conds 1024 --reps 1"
$CMD > gpurun_out/plain_enc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_encoder_umma -s 1 -c 1 -f -o gpurun_out/prof_enc $CMD > gpurun_out/ncu_enc.log 2>&1
tail -n 3 gpurun_out/plain_enc.log; tail -n 3 gpurun_out/ncu_enc.log
