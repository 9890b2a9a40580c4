mkdir -p gpurun_out
for cfg in "37 p,8" "36 p,8" "74 p,2"; do
set -- $cfg
JJ_BENCH_NPARTS=$1 JJ_SUBDOMAIN=$2 JJ_SUB_PROF=1 JJ_BENCH_INNER=100 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/itp.json 2> gpurun_out/itp.err
echo "=== n_parts $1 cfg $2"; grep -A 40 "JJ_SUB_PROF" gpurun_out/itp.err | tail -22 | grep -v "sweep level\|local cycles"; tail -2 gpurun_out/itp.err | cut -c1-200
done
