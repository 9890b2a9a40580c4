"""cProfile of the second compute() call of a named configuration (host side of the end-to-end figure).
    python tools/e2e_profile_cfg.py cfg4 [time steps]"""
import cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyjjasim_b200 as pj
from pyjjasim_b200 import engine
import bench
name = sys.argv[1]
a, kw, Nt, (w0, w1), W_total, note = bench.named_config(pj, name, 0, 1)
if len(sys.argv) > 2:
    Nt = int(sys.argv[2])
def one():
    prob = pj.TimeEvolutionProblem(a, time_step_count=Nt, store_time_steps=[Nt - 1], store_current=False, store_voltage=False, **kw)
    return prob.compute()
t = time.perf_counter(); one(); print("first call %.3f s" % (time.perf_counter() - t))
t = time.perf_counter(); one(); print("second call %.3f s" % (time.perf_counter() - t))
pr = cProfile.Profile(); pr.enable(); one(); pr.disable()
st = list(engine.last_run_stats.values())[0]
print("device ms", st["total_ms"])
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
