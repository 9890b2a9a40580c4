"""Per-launch cost of the step kernel: device time of runs of n steps, n = 1, 10, 100 (annealing intervals are 10 steps)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyjjasim_b200 as pj
from pyjjasim_b200 import engine as eng_mod, _lib

a = pj.SquareArray(100, 100)
W, dt = 256, 0.5
prob = pj.TimeEvolutionProblem(a, time_step=dt, time_step_count=10, external_flux=0.1, current_sources=0,
                               temperature=np.full((1, W, 1), 0.3), store_current=False, store_voltage=False, noise_seed=1)
tab = eng_mod._tables_for(a, dt, n_parts=eng_mod._n_parts_for(a, W, 0, None))
key, e, kind = eng_mod._engine_for(tab, pj.DefaultCPR(), 0, W, eng_mod._engine_kind(None))
e.set_problem(W, dt, seed=1, engine=kind)
e.alloc_outputs(10, 0)
specs = eng_mod._classify_all(prob, tab)
eng_mod._setup_sources(e, specs, eng_mod._ShardInputs(specs, 0, W), tab)
out = {}
i0 = 0
for n, reps, planes in ((10, 3, None), (1, 20, None), (10, 20, None), (100, 3, None), (10, 20, np.arange(10))):
    ms = []
    for r in range(reps):
        e.run(i0, n, planes, None)
        ms.append(e.stats()["step_ms"])
        i0 += n
    out["n=%d%s" % (n, " +planes" if planes is not None else "")] = dict(ms_per_launch=float(np.median(ms)), us_per_step=float(np.median(ms)) * 1e3 / n)
print(json.dumps(out, indent=1))
