# Round-2 closing single-GPU pass (after the split-precision chain and the 4-instruction KDE term): the whole test suite,
# smoke, the sweeps / benches and the captures that profiles/r02_* record.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest-all rc=$?"; tail -n 3 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 6 gpurun_out/smoke.log
timeout 300 python scripts/chain_sweep.py --T 200 --members 1024,4096,18944,37888 --precisions fp32,bf16,bf16x3 --reps 2 2>&1 | tee gpurun_out/chain_sweep_h.log
timeout 300 python scripts/chain_sweep.py --T 1000 --members 1024,18944 --precisions bf16,bf16x3 --reps 1 2>&1 | tee -a gpurun_out/chain_sweep_h.log
timeout 300 python scripts/measure_parity.py > gpurun_out/parity.log 2>&1; echo "parity rc=$?"
bash scripts/gpu_umma_timing.sh > /dev/null 2>&1; tail -n 8 gpurun_out/sweep_timing.log | cut -c1-300
timeout 300 python scripts/summary_window_bench.py --chain --cases 151552x4,8192x4,2048x4,18944x29,8192x29,256x29 2>/dev/null | tee gpurun_out/summary_window_chain.log
timeout 600 python scripts/stats_bench.py > gpurun_out/stats_bench.log 2>&1; echo "stats_bench rc=$?"; grep KDE gpurun_out/stats_bench.log
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_r02_final.err
A="--steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs"
timeout 300 python bench.py $A > gpurun_out/plain_a.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_fp32_b256.csv python bench.py $A > gpurun_out/ncu_a.log 2>&1
B="$A --precision bf16 --members 18944"
timeout 300 python bench.py $B > gpurun_out/plain_b.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bf16_b18944.csv python bench.py $B > gpurun_out/ncu_b.log 2>&1
D="python scripts/chain_sweep.py --members 18944 --precisions bf16x3 --T 200 --reps 1"
timeout 300 $D > gpurun_out/plain_d.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_umma -s 1 -c 1 -f -o gpurun_out/prof_chain_umma_split $D > gpurun_out/ncu_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_kde_scan32 -c 1 -f -o gpurun_out/prof_kde_scan python scripts/stats_bench.py --maps 1024 --fields "" --reps 0 --only kde > gpurun_out/ncu_s3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
