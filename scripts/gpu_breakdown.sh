python scripts/step_breakdown.py
