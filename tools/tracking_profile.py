"""cProfile of one warm compute() of the cfg3 vortex-tracking workload with running observables (mode c of
vortex_tracking_bench.py): where the host time beside the device run goes."""
import cProfile, pstats, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyjjasim_b200 as pj
Nt, every, W = 1000, 100, 512
a = pj.HoneycombArray(200, 200)
Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0.1, 1.5, W))
kw = dict(circuit=a, time_step=0.05, time_step_count=Nt, external_flux=0.1, current_sources=Is,
          store_time_steps=np.arange(every - 1, Nt, every), store_current=False, store_voltage=False,
          store_theta=False, observe_interval=every, observe_first=every - 1)
pj.TimeEvolutionProblem(**kw).compute()
pr = cProfile.Profile(); pr.enable()
res = pj.TimeEvolutionProblem(**kw).compute(); nsum = res.get_vortex_sum()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
