for m in 512 1024 2048; do echo fused; python scripts/step_breakdown.py --members $m | grep kde; echo staged; ERTDIFF_KDE_SMALL_N=500 python scripts/step_breakdown.py --members $m | grep kde; done
