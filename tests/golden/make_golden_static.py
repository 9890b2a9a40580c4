"""Golden vectors for the static polish (pyjjasim_b200/static_polish.py) from the UNMODIFIED reference:
StaticProblem.approximate() and StaticProblem.compute() (static_problem.py:550-610) on a few vortex configurations.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_static.py"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_harness import import_reference  # noqa: E402
from tests.cases import static_cases              # noqa: E402

ref = import_reference()
for name, (ctor, args, L, f, n, Is) in static_cases(ref).items():
    a = getattr(ref, ctor)(*args)
    if L:
        a.set_inductance(L)
    W = n.shape[1]
    th0, th, status, iters = [], [], [], []
    for w in range(W):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sp = ref.StaticProblem(a, current_sources=Is[:, w].copy(), external_flux=f[:, w].copy(),
                                   vortex_configuration=n[:, w].copy())
            th0.append(sp.approximate()._th())
            cfg, st, info = sp.compute()
        th.append(cfg._th()); status.append(st); iters.append(info.get_number_of_iterations())
    out = os.path.join(ROOT, "tests", "golden", f"static_{name}.npz")
    np.savez_compressed(out, theta0=np.stack(th0, 1), theta=np.stack(th, 1), status=np.array(status), iterations=np.array(iters))
    print(name, "status", status, "iterations", iters, "->", out)
