# Round-2 final single-GPU pass: the whole test suite, smoke, the benches and the launch lists that profiles/r02_* record.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest-all rc=$?"; tail -n 3 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 300 python scripts/summary_window_bench.py --chain --cases 151552x4,8192x4,2048x4,18944x29,8192x29,256x29 2>/dev/null | tee gpurun_out/summary_window_chain.log
timeout 600 python scripts/stats_bench.py > gpurun_out/stats_bench.log 2>&1; echo "stats_bench rc=$?"; cat gpurun_out/stats_bench.log
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_r02_final.err
A="--steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs"
timeout 300 python bench.py $A > gpurun_out/plain_a.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_fp32_b256.csv python bench.py $A > gpurun_out/ncu_a.log 2>&1
B="$A --precision bf16 --members 18944"
timeout 300 python bench.py $B > gpurun_out/plain_b.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bf16_b18944.csv python bench.py $B > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_percentiles_select -c 1 -f -o gpurun_out/prof_percentiles python scripts/stats_bench.py --maps 1024 --fields "" --reps 0 --only pct > gpurun_out/ncu_s2.log 2>&1
D="python scripts/chain_sweep.py --members 18944 --precisions bf16 --T 200 --reps 1"
timeout 300 $D > gpurun_out/plain_d.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_umma -s 1 -c 1 -f -o gpurun_out/prof_chain_umma $D > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
