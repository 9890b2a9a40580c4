mkdir -p gpurun_out
for tt in 1024 2048 4096 8192; do
JJ_TT_MAX=$tt timeout 900 python tools/config_sweep.py cfg5 > gpurun_out/r2_tt_$tt.jsonl 2> gpurun_out/r2_tt_$tt.err
echo "tt_max=$tt $(cut -c1-330 gpurun_out/r2_tt_$tt.jsonl)"; tail -2 gpurun_out/r2_tt_$tt.err | cut -c1-200
done
for tt in 2048 4096; do
JJ_TT_MAX=$tt timeout 900 python tools/config_sweep.py cfg4 cfg3 > gpurun_out/r2_tt34_$tt.jsonl 2> gpurun_out/r2_tt34_$tt.err
echo "tt_max=$tt"; cut -c1-330 gpurun_out/r2_tt34_$tt.jsonl
done
