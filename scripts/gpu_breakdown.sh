python scripts/step_breakdown.py --members 2048 | grep -v run_chain
python scripts/step_breakdown.py --members 1024 | grep -v run_chain
python scripts/step_breakdown.py --members 512 | grep -v run_chain
