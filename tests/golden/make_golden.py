"""
Generate golden vectors for the time-evolution hot path by running the UNMODIFIED reference
(/root/reference, importable only in the build container) on the parity cases of tests/cases.py.

    python tests/golden/make_golden.py

Writes tests/golden/<case>.npz with the reference's stored theta / current / voltage arrays
((Nj, W, Nt_s) float64), a checksum of the circuit's cycle matrix and the numpy/scipy versions used.
For T > 0 cases the global numpy generator is seeded with the case seed right before compute(), so the
draw sequence is the one tests/cases.replay_noise reproduces.
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness import import_reference  # noqa: E402
from tests import cases  # noqa: E402


def matrix_digest(A):
    A = A.tocsr()
    A.sort_indices()
    h = hashlib.sha256()
    for arr in (A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.int64)):
        h.update(arr.tobytes())
    return h.hexdigest()


def main():
    ref = import_reference()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name in cases.CASES:
        kw, seed = cases.build(name, ref)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            prob = ref.TimeEvolutionProblem(**kw)
            if seed is not None:
                np.random.seed(seed)
            res = prob.compute()
        data = dict(A_digest=matrix_digest(kw["circuit"].get_cycle_matrix()),
                    Nj=kw["circuit"]._Nj(), Nf=kw["circuit"]._Nf(), W=prob.get_problem_count(),
                    numpy_version=np.__version__, scipy_version=scipy.__version__)
        for key in ("theta", "current", "voltage"):
            arr = getattr(res, key)
            if arr is not None:
                data[key] = arr
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **data)
        print(name, {k: getattr(v, "shape", v) for k, v in data.items()})


if __name__ == "__main__":
    main()
