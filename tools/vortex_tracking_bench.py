"""cfg3 (HoneycombArray(200,200), f = 0.1, DC bias sweep, 512 problems) with vortex tracking every 100 steps, three ways:
(a) phase planes stored and copied to the host, vortex configurations derived there (the reference's way);
(b) store_vortex_configuration=True, store_theta=False: n computed on the device from the stored planes, copied as int32;
(c) running observables only: vortex sums and DC voltages accumulated inside the step kernel, nothing stored.
Prints one JSON line (wall seconds of a warm compute() call, junction-steps/s end to end, bytes copied to the host)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyjjasim_b200 as pj  # noqa: E402
from pyjjasim_b200 import engine  # noqa: E402

Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
every = 100
a = pj.HoneycombArray(200, 200)
W = 512
Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0.1, 1.5, W))
base = dict(circuit=a, time_step=0.05, time_step_count=Nt, external_flux=0.1, current_sources=Is,
            store_time_steps=np.arange(every - 1, Nt, every), store_current=False, store_voltage=False)
modes = {
    "a_theta_planes_to_host": dict(store_theta=True),
    "b_vortex_planes_on_device": dict(store_theta=False, store_vortex_configuration=True),
    "c_running_observables": dict(store_theta=False, observe_interval=every, observe_first=every - 1),
}
out = dict(workload=f"cfg3 HoneycombArray(200,200) 512 problems, {Nt} steps, tracking every {every} steps", Nj=a._Nj(), Nf=a._Nf())
ref_n = None
for name, extra in modes.items():
    kw = dict(base, **extra)
    pj.TimeEvolutionProblem(**kw).compute()                 # setup + pinning
    t0 = time.perf_counter()
    res = pj.TimeEvolutionProblem(**kw).compute()
    if name.startswith("a"):
        n = res.get_vortex_configuration()
        copied = res.theta.nbytes
    elif name.startswith("b"):
        n = res.get_vortex_configuration()
        copied = n.shape[0] * n.shape[1] * n.shape[2] * 4
    else:
        nsum = res.get_vortex_sum()
        copied = nsum.size * 4 + 2 * a._Nj() * W * 8
    wall = time.perf_counter() - t0
    st = list(engine.last_run_stats.values())[0]
    if name.startswith("a"):
        ref_n = n
    elif name.startswith("b"):
        assert np.array_equal(n, ref_n)
    else:
        assert np.array_equal(nsum, ref_n.sum(axis=2))
    out[name] = dict(wall_s=round(wall, 3), device_s=round(st["total_ms"] * 1e-3, 3), host_bytes=int(copied),
                     junction_steps_per_s_e2e=a._Nj() * W * Nt / wall)
    del res
print(json.dumps(out))
