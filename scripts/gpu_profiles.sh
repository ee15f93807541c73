# Round-end evidence: launch lists (shares) and one `ncu --set full` capture per hot kernel.
# Each command runs plain first (must exit 0), then under ncu; numbers printed under ncu are never bench values.
# usage: gpu_profiles.sh launches|chain_fp32|chain_umma|chain_umma2|encoder_umma   (one .ncu-rep per call: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
case "$1" in
launches)
A="--steps 3 --warmup 3 --no-cpu-baseline"
python bench.py $A > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_fp32_b256.csv python bench.py $A > gpurun_out/ncu_a.log 2>&1
B="--steps 3 --warmup 3 --no-cpu-baseline --precision bf16 --members 8192"
python bench.py $B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bf16_b8192.csv python bench.py $B > gpurun_out/ncu_b.log 2>&1
;;
chain_fp32)
C="python scripts/chain_sweep.py --members 256 --precisions fp32 --T 1000 --reps 1"
$C > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_chain -s 1 -c 1 -f -o gpurun_out/prof_chain_fp32 $C > gpurun_out/ncu_c.log 2>&1
;;
chain_umma)
D="python scripts/chain_sweep.py --members 18944 --precisions bf16 --T 200 --reps 1"
$D > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_chain_umma -s 1 -c 1 -f -o gpurun_out/prof_chain_umma $D > gpurun_out/ncu_d.log 2>&1
;;
chain_umma2)
F="python scripts/chain_sweep.py --members 37888 --precisions bf16 --T 200 --reps 1"
$F > gpurun_out/plain_f.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_chain_umma -s 1 -c 1 -f -o gpurun_out/prof_chain_umma2 $F > gpurun_out/ncu_f.log 2>&1
;;
encoder_umma)
E="python scripts/encoder_bench.py --conds 1024 --reps 1"
$E > gpurun_out/plain_e.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_encoder_umma -s 1 -c 1 -f -o gpurun_out/prof_encoder_umma $E > gpurun_out/ncu_e.log 2>&1
;;
esac
ls -la gpurun_out/*.ncu-rep 2>/dev/null
