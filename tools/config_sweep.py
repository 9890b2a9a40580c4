"""
The BASELINE.json configurations other than the bench workload, run through the public API on one GPU with a short
horizon: which engine takes them, host setup time, device time per time step and junction-steps/s. cfg4 runs one GPU's
share of its 4096 problems (512). Prints one JSON line per configuration.

    python tools/config_sweep.py [cfg1 cfg3 cfg4 cfg5]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyjjasim_b200 as pj  # noqa: E402
from pyjjasim_b200 import engine  # noqa: E402


def cfg1():
    a = pj.SquareArray(20, 20)
    Is = a.current_base(angle=0)[:, None, None] * np.linspace(0, 2, 32)[None, :, None]
    return a, dict(time_step=0.05, time_step_count=10000, current_sources=Is)


def cfg3():
    a = pj.HoneycombArray(200, 200)
    Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0.1, 1.5, 512))
    return a, dict(time_step=0.05, time_step_count=40, external_flux=0.1, current_sources=Is)


def cfg4():
    a = pj.SquareArray(256, 256)
    a.set_capacitance(1.0)
    W = 512
    IDC, IA = np.linspace(0, 2, W), np.linspace(0, 3, W)
    Is = pj.RankOneSource(a.current_base(angle=0), lambda i: IDC + IA * np.sin(0.25 * i * 0.05), problem_count=W)
    return a, dict(time_step=0.05, time_step_count=40, current_sources=Is, temperature=0.01 * np.ones((1, W, 1)), noise_seed=1234)


def cfg5():
    a = pj.SquareArray(1000, 1000)
    a.set_inductance(1.0)
    a.set_capacitance(1.0)
    Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0.5, 1.5, 64))
    return a, dict(time_step=0.05, time_step_count=10, external_flux=0.05, current_sources=Is)


def main():
    names = sys.argv[1:] or ["cfg1", "cfg3", "cfg4"]
    for name in names:
        t0 = time.perf_counter()
        a, kw = globals()[name]()
        Nt = kw["time_step_count"]
        prob = pj.TimeEvolutionProblem(a, store_time_steps=[Nt - 1], store_current=False, store_voltage=False, **kw)
        t1 = time.perf_counter()
        res = prob.compute()                   # first call: ordering, factorisation, plans, upload
        t2 = time.perf_counter()
        res = prob.compute()
        t3 = time.perf_counter()
        st = list(engine.last_run_stats.values())[0]
        W = prob.get_problem_count()
        out = dict(config=name, Nj=a._Nj(), Nf=a._Nf(), W=W, time_steps=Nt,
                   engine={1: "streaming", 3: "subdomain"}.get(st["engine"]),
                   setup_s=round(t2 - t1 - (t3 - t2), 2), device_us_per_time_step=round(st["total_ms"] * 1e3 / Nt, 1),
                   junction_steps_per_s_device=a._Nj() * W * Nt / (st["total_ms"] * 1e-3),
                   junction_steps_per_s_e2e=a._Nj() * W * Nt / (t3 - t2), finite=bool(np.all(np.isfinite(res.theta))),
                   device_MB=round(st["device_bytes"] / 1e6, 1))
        if os.environ.get("SWEEP_CHECK") and out["engine"] != "streaming":
            # same problem on the streaming engine (same Philox counters): the two engines must agree
            os.environ["JJ_ENGINE"] = "streaming"
            try:
                ref = prob.compute()
            finally:
                os.environ.pop("JJ_ENGINE", None)
            out["max_abs_dtheta_vs_streaming"] = float(np.max(np.abs(ref.theta - res.theta)))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
