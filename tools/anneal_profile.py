import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyjjasim_b200 as pj
a = pj.SquareArray(100, 100)
kw = dict(circuit=a, time_step=0.5, interval_steps=10, external_flux=0.1, problem_count=256,
          interval_count=200, vortex_mobility=0.001, start_T=0.3, T_factor=1.03, noise_seed=1234)
pj.AnnealingProblem(**{**kw, "interval_count": 3}).anneal()
pr = cProfile.Profile(); pr.enable(); pj.AnnealingProblem(**kw).anneal(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
