timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/e2e_profile.py 2>&1 | head -22
