mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "anneal or mobility or phase_zones or observ or vortex" > gpurun_out/r2_pytest_anneal.txt 2>&1
tail -5 gpurun_out/r2_pytest_anneal.txt
ANNEAL_LOOP=0 timeout 600 python tools/anneal_bench.py 2>/dev/null | head -c 420; echo
ANNEAL_LOOP=0 timeout 600 python tools/anneal_bench.py 2>/dev/null | head -c 420; echo
JJ_ANNEAL_ZONES=0 ANNEAL_LOOP=0 timeout 600 python tools/anneal_bench.py 2>/dev/null | head -c 420; echo
