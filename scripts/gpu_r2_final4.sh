# the two captures that changed with the coarse-to-fine KDE scan: its ncu --set full capture and the bf16 18,944-member launch list
mkdir -p gpurun_out
timeout 80 ncu --set full --clock-control none --import-source on -k regex:k_kde_scan32 -c 1 -f -o gpurun_out/prof_kde_scan python scripts/stats_bench.py --maps 1024 --fields "" --reps 0 --only kde > gpurun_out/ncu_s3.log 2>&1; echo "ncu kde rc=$?"
B="--steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs --precision bf16 --members 18944"
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bf16_b18944.csv python bench.py $B > gpurun_out/ncu_b.log 2>&1; echo "launch list rc=$?"
