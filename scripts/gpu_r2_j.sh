# Round-2 pass J: the coarse-to-fine KDE scan -- statistics tests, then A/B against the full scan (ERTDIFF_KDE_COARSE=1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stats.py -m gpu -q -x > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_j.log
for c in 1 32; do
  echo "== ERTDIFF_KDE_COARSE=$c"
  ERTDIFF_KDE_COARSE=$c timeout 300 python scripts/stats_bench.py --only kde 2>&1 | grep KDE
  ERTDIFF_KDE_COARSE=$c timeout 300 python scripts/summary_window_bench.py --chain --cases 151552x4,8192x4,2048x4,18944x29,8192x29 2>/dev/null | grep members | cut -c1-260
done | tee gpurun_out/kde_coarse_ab.log
