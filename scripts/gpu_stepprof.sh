MEMBERS=8192 PRECISION=bf16 python scripts/step_profile.py 2>&1 | grep "ertdiff\|sum per"
