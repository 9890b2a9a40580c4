"""A three-step run of the bench workload (cfg2) and of a circuit with upper phases, for compute-sanitizer
(memcheck / synccheck / racecheck logs kept under profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyjjasim_b200 as pj
from pyjjasim_b200 import engine
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
if which == "cfg2":
    a = pj.SquareArray(100, 100); W = 256
    kw = dict(time_step=0.5, external_flux=0.1, temperature=np.geomspace(1e-2, 1, W)[None, :, None], noise_seed=1)
elif which == "anneal":
    # the annealing schedule: zone-byte variant of the step kernel, k_zone_mobility, amplitude / rule kernels, hand-over
    # of the state to the half-step engine
    a = pj.SquareArray(40, 37)
    th, n, prof = pj.AnnealingProblem(a, time_step=0.5, interval_steps=5, external_flux=0.2, problem_count=40, interval_count=4,
                                      vortex_mobility=0.02, start_T=0.4, T_factor=1.25, noise_seed=3).anneal()
    print(which, "engine", engine.last_run_stats[0]["engine"], "finite", bool(np.all(np.isfinite(th))), "T", float(prof[-1].mean()))
    sys.exit(0)
else:
    # several items per block + upper program (the cfg3 / cfg4 / cfg5 regime, scaled down)
    os.environ["JJ_TT_MAX"] = "120"
    a = pj.SquareArray(80, 80); W = 512
    kw = dict(time_step=0.05, external_flux=0.1, current_sources=pj.RankOneSource(a.current_base(angle=0), np.linspace(0.2, 1.8, W)))
res = pj.TimeEvolutionProblem(a, time_step_count=3, store_time_steps=[2], store_voltage=False, **kw).compute()
print(which, "engine", engine.last_run_stats[0]["engine"], "finite", bool(np.all(np.isfinite(res.theta))))
