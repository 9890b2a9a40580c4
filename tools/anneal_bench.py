"""
Annealing caller of the path measured end to end on one GPU:
  device   AnnealingProblem.anneal()  - schedule resident on the GPU (state, factor, theta planes in HBM)
  loop     the reference's loop written against TimeEvolutionProblem.compute() (state and theta planes cross PCIe
           every interval, vortex configurations on the host) - what a drop-in user gets without the resident path
  cpu      the numpy/scipy oracle restatement of AnnealingProblem.compute on a bounded sample, scaled linearly
Prints one JSON line; junction-steps/s = Nj * W * (interval_count * interval_steps + 5 * interval_steps) / wall.
"""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyjjasim_b200 as pj  # noqa: E402


def main():
    N = int(os.environ.get("ANNEAL_N", "100"))
    W = int(os.environ.get("ANNEAL_W", "256"))
    count = int(os.environ.get("ANNEAL_INTERVALS", "200"))
    steps = 10
    a = pj.SquareArray(N, N)
    kw = dict(circuit=a, time_step=0.5, interval_steps=steps, external_flux=0.1, problem_count=W,
              interval_count=count, vortex_mobility=0.001, start_T=0.3, T_factor=1.03, noise_seed=1234)
    Nj = a._Nj()
    total = Nj * W * (count + 5) * steps
    out = dict(workload=f"AnnealingProblem SquareArray({N},{N}) f=0.1, {W} problems, {count} intervals x {steps} steps + 5 closing runs",
               Nj=Nj, W=W)
    pj.AnnealingProblem(**{**kw, "interval_count": 3}).anneal()          # warm-up: tables, plans, engines
    t0 = time.perf_counter()
    ap = pj.AnnealingProblem(**kw)
    theta, n, prof = ap.anneal()
    t1 = time.perf_counter()
    dev_ms = sum(s["total_ms"] for s in ap.last_stats.values())
    out["device"] = dict(wall_s=t1 - t0, device_ms=dev_ms, junction_steps_per_s=total / (t1 - t0),
                         final_T_median=float(np.median(prof[-1])), vortices_mean=float(np.abs(n).sum(axis=0).mean()))
    if os.environ.get("ANNEAL_LOOP", "1") == "0":
        print(json.dumps(out))
        return
    # reference-style loop through compute()
    cnt2 = max(3, count // 10)
    f = np.atleast_1d(0.1)[:, None, None]
    rs = pj.AnnealingProblem(**{**kw, "interval_count": cnt2})
    th = np.zeros((Nj, W))
    prob = pj.TimeEvolutionProblem(a, time_step_count=steps, time_step=0.5, external_flux=f, current_sources=0,
                                   temperature=rs.T, store_current=False, store_voltage=False, noise_seed=1234)
    t0 = time.perf_counter()
    for i in range(cnt2):
        prob.temperature = rs.T * np.ones((1, 1, steps))
        prob.config_at_minus_1 = th
        prob.config_at_minus_2 = th.copy()
        res = prob.compute()
        rs._temperature_adjustment(rs.get_vortex_mobility(res.get_vortex_configuration()), i)
        th = res.get_theta()[..., -1]
    t1 = time.perf_counter()
    out["loop"] = dict(intervals=cnt2, wall_s=t1 - t0, junction_steps_per_s=Nj * W * cnt2 * steps / (t1 - t0))
    if os.environ.get("ANNEAL_CPU", "1") != "0":
        from oracle import oracle
        Wc, cc = min(W, 32), 2
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            oracle.annealing(a.get_cycle_matrix(), a._Ic(), a._R(), a._C(), a._L(), time_step=0.5, interval_steps=steps,
                             external_flux=0.1, problem_count=Wc, interval_count=cc, start_T=0.3,
                             rng=np.random.RandomState(0), final_runs=0)
        t1 = time.perf_counter()
        out["cpu"] = dict(kind="port", sample=f"{Wc} problems x {cc} intervals, no closing runs", wall_s=t1 - t0,
                          junction_steps_per_s=Nj * Wc * cc * steps / (t1 - t0), cores=os.cpu_count())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
