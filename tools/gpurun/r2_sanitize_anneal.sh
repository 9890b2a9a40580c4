mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py anneal > gpurun_out/sanitizer_memcheck_anneal.txt 2>&1
tail -3 gpurun_out/sanitizer_memcheck_anneal.txt
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_run.py anneal > gpurun_out/sanitizer_synccheck_anneal.txt 2>&1
tail -3 gpurun_out/sanitizer_synccheck_anneal.txt
