timeout 900 python -m pytest tests -m gpu -x -q -k "pinned_pool" 2>&1 | tail -12
