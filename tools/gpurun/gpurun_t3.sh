timeout 900 python -m pytest tests -m gpu -x -q -k "long_noisy" 2>&1 | tail -25
