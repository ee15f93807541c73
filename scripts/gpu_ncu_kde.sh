mkdir -p gpurun_out
CMD="python scripts/kde_only.py"
$CMD > gpurun_out/plain_k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_kde_small -s 1 -c 1 -f -o gpurun_out/prof_kde_small $CMD > gpurun_out/ncu_k.log 2>&1
tail -n 2 gpurun_out/ncu_k.log
