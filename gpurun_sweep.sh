mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "resident" 2>&1 | tail -4
for cfg in 4,8 8,8; do
for dbg in 0 2; do
  JJ_RES_DEBUG=$dbg JJ_RESIDENT=$cfg JJ_BENCH_INNER=500 JJ_BENCH_SKIP_E2E=1 python bench.py --steps 2 --warmup 1 > gpurun_out/dbg_$dbg.json 2> gpurun_out/dbg_$dbg.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dbg_$dbg.json"))
    print("cfg $cfg skipmask $dbg", "us/timestep %.1f"%(d["ms_per_step"]*1e3/d["config"]["time_steps_per_step"]))
except Exception as e:
    print("$dbg failed", e, open("gpurun_out/dbg_$dbg.err").read()[-300:])
PY
done
done
