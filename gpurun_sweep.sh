mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "resident" 2>&1 | tail -4
for leaf in 8 16 32; do
for cfg in 4,8 8,8; do
  JJ_LEAF_SIZE=$leaf JJ_RESIDENT=$cfg JJ_BENCH_INNER=500 JJ_BENCH_SKIP_E2E=1 python bench.py --steps 3 --warmup 1 > gpurun_out/sweep_$cfg.json 2> gpurun_out/sweep_$cfg.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$cfg.json"))
    print("leaf $leaf cfg $cfg", "%.3e js/s"%d["value"], "us/timestep %.1f"%(d["ms_per_step"]*1e3/d["config"]["time_steps_per_step"]), "frac %.3f"%d["roofline"]["frac"], d["config"]["engine"])
except Exception as e:
    print("$cfg failed", e, open("gpurun_out/sweep_$cfg.err").read()[-800:])
PY
done
done
