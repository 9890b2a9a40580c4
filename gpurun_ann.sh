set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "anneal or observables" 2>&1 | tail -25
