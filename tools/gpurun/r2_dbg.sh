mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "subdomain_engine_matches_reference_golden or multiple_items" 2>&1 | grep -v "^tests.*PASSED" | grep "Error\|error\|assert\|FAILED\|passed\|failed" | head -60 > gpurun_out/r2_dbg.txt
cat gpurun_out/r2_dbg.txt
