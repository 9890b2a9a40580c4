import cProfile, pstats, time, os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import pyjjasim_b200 as pj
a = pj.SquareArray(100, 100)
W = 256; INNER = 1000
T = np.geomspace(1e-2, 1.0, W)
th0 = np.zeros((a._Nj(), W))
def one():
    global th0
    prob = pj.TimeEvolutionProblem(a, time_step=0.5, time_step_count=INNER, external_flux=0.1,
                                   temperature=T[None, :, None], store_time_steps=[INNER // 3, INNER - 1],
                                   store_current=False, store_voltage=False, config_at_minus_1=th0, noise_seed=1234)
    res = prob.compute()
    th0 = np.ascontiguousarray(res.theta[:, :, -1])
t = time.perf_counter(); one(); print("first call %.3f s" % (time.perf_counter() - t))
t = time.perf_counter(); one(); print("second call %.3f s" % (time.perf_counter() - t))
pr = cProfile.Profile(); pr.enable(); one(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
