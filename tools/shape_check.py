"""Robustness sweep: assorted circuit shapes and problem counts through the auto engine against the streaming engine
(same Philox counters), a few time steps each. Prints engine, layout and max |dtheta|."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyjjasim_b200 as pj  # noqa: E402
from pyjjasim_b200 import engine  # noqa: E402

CASES = [("SquareArray", (150, 150), 64), ("SquareArray", (200, 120), 200), ("HoneycombArray", (60, 60), 300),
         ("SquareArray", (33, 77), 19), ("TriangularArray", (40, 30), 100), ("SquareArray", (120, 120), 1000),
         ("HoneycombArray", (24, 90), 36), ("SquareArray", (64, 64), 7),
         # more items than blocks with few problems (work-counter scheduling, z through TMA bulk copies), ragged widths
         ("SquareArray", (300, 300), 20), ("HoneycombArray", (150, 100), 52), ("SquareArray", (90, 310), 130),
         ("SquareArray", (400, 50), 3), ("SquareArray", (6, 5), 2), ("SquareArray", (12, 12), 1500)]


def main():
    worst = 0.0
    for kind, shape, W in CASES:
        a = getattr(pj, kind)(*shape)
        if kind == "SquareArray" and shape[0] % 2 == 0:
            a.set_capacitance(0.5)
        Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0.3, 1.6, W))
        kw = dict(circuit=a, time_step=0.05, time_step_count=8, external_flux=0.13, current_sources=Is,
                  temperature=0.05 * np.ones((1, W, 1)), noise_seed=11, store_time_steps=[7],
                  store_current=False, store_voltage=False)
        # (running observables on: the vortex pass runs in every kernel variant)
        res = pj.TimeEvolutionProblem(observe_interval=3, observe_first=1, **kw).compute()
        st = list(engine.last_run_stats.values())[0]
        os.environ["JJ_ENGINE"] = "streaming"
        try:
            ref = pj.TimeEvolutionProblem(**kw).compute()
        finally:
            os.environ.pop("JJ_ENGINE", None)
        obs = pj.TimeEvolutionProblem(observe_interval=3, observe_first=1, store_time_steps=[1, 4, 7], **{k: v for k, v in kw.items() if k != "store_time_steps"}).compute()
        n_sum = obs.get_vortex_configuration().sum(axis=2)
        assert np.array_equal(res.get_vortex_sum(), n_sum), "vortex sums"
        err = float(np.max(np.abs(res.theta - ref.theta)))
        engine._tables_cache.clear()
        worst = max(worst, err)
        print(f"{kind}{shape} W={W}: Nf={a._Nf()} engine={st['engine']} subdomains={st['cluster_size']} "
              f"problems/block={st['tile_problems']} max|dtheta|={err:.2e}", flush=True)
    print("worst", worst)
    assert worst < 1e-9


if __name__ == "__main__":
    main()
