"""Short annealing schedule for an ncu launch list (per-kernel durations of one interval)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyjjasim_b200 as pj
a = pj.SquareArray(100, 100)
n = int(os.environ.get("JJ_ANNEAL_INTERVALS", "6"))
pj.AnnealingProblem(circuit=a, time_step=0.5, interval_steps=10, external_flux=0.1, problem_count=256,
                    interval_count=n, vortex_mobility=0.001, start_T=0.3, T_factor=1.03, noise_seed=1234).anneal()
