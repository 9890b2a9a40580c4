mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "anneal or mobility or zone" > gpurun_out/r2_pytest_anneal.txt 2>&1
tail -3 gpurun_out/r2_pytest_anneal.txt
ANNEAL_LOOP=0 timeout 600 python tools/anneal_bench.py 2>/dev/null | head -c 420; echo
ANNEAL_LOOP=0 timeout 600 python tools/anneal_bench.py 2>/dev/null | head -c 420; echo
timeout 300 python tools/anneal_profile.py 2>&1 | head -16 | tail -9
