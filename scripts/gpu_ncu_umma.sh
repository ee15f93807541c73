mkdir -p gpurun_out
CMD="python scripts/chain_sweep.py --members 18944 --precisions bf16 --T 200 --reps 1"
$CMD > gpurun_out/plain_umma.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_chain_umma -s 1 -c 1 -f -o gpurun_out/prof_umma4 $CMD > gpurun_out/ncu_umma.log 2>&1
tail -n 3 gpurun_out/plain_umma.log; tail -n 3 gpurun_out/ncu_umma.log
