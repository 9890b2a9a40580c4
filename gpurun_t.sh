timeout 900 python -m pytest tests -m gpu -x -q -k "larger_circuit" 2>&1 | tail -5
