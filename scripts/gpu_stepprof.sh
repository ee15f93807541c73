for g in 4 8 16 32; do echo "gchunks $g"; ERTDIFF_KDE_GCHUNKS=$g python scripts/step_profile.py 2>&1 | grep "k_kde_small"; done
echo default; python scripts/step_profile.py 2>&1 | grep "k_kde_small\|sum per"
