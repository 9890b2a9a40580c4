import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, pyjjasim_b200 as pj
from pyjjasim_b200 import engine
a = pj.SquareArray(1000, 1000); a.set_inductance(1.0); a.set_capacitance(1.0)
for W in (8, 16):
    Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0.5, 1.5, 64)[:W])
    t0 = time.time()
    prob = pj.TimeEvolutionProblem(a, time_step=0.05, time_step_count=12, external_flux=0.05, current_sources=Is,
                                   store_time_steps=[11], store_current=False, store_voltage=False)
    res = prob.compute(); t1 = time.time(); res = prob.compute(); t2 = time.time()
    st = list(engine.last_run_stats.values())[0]
    print("W", W, "engine", st["engine"], "subdomains", st["cluster_size"], "PC", st["tile_problems"], "setup %.1f s" % (t1 - t0 - (t2 - t1)),
          "us/step %.1f" % (st["total_ms"] * 1e3 / 12), "G js/s %.2f" % (a._Nj() * W * 12 / (st["total_ms"] * 1e-3) / 1e9), "finite", bool(np.all(np.isfinite(res.theta))), flush=True)
    engine._tables_cache.clear()
