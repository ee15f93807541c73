# Round-2 second GPU pass: full test suite with the measured tolerances, wider tiling sweep, statistics bench,
# launch lists and ncu captures.  Every step under `timeout`; plain run first, ncu second.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest-all rc=$?"; tail -n 8 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 300 python scripts/chain_fp32_variants.py --members 444,592,700,1184,2048,4096,8192 --reps 2 > gpurun_out/variants_h128_b.log 2>&1; cat gpurun_out/variants_h128_b.log
timeout 600 python scripts/stats_bench.py > gpurun_out/stats_bench.log 2>&1; echo "stats_bench rc=$?"; cat gpurun_out/stats_bench.log
A="--steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs"
timeout 300 python bench.py $A > gpurun_out/plain_a.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_fp32_b256.csv python bench.py $A > gpurun_out/ncu_a.log 2>&1
B="$A --precision bf16 --members 18944"
timeout 300 python bench.py $B > gpurun_out/plain_b.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bf16_b18944.csv python bench.py $B > gpurun_out/ncu_b.log 2>&1
B2="$A --precision bf16 --members 8192"
timeout 300 python bench.py $B2 > gpurun_out/plain_b2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bf16_b8192.csv python bench.py $B2 > gpurun_out/ncu_b2.log 2>&1
C="python scripts/chain_sweep.py --members 256 --precisions fp32 --T 1000 --reps 1"
timeout 300 $C > gpurun_out/plain_c.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 1 -c 1 -f -o gpurun_out/prof_chain_fp32 $C > gpurun_out/ncu_c.log 2>&1
cat gpurun_out/plain_c.log
ls -la gpurun_out/*.ncu-rep 2>/dev/null
