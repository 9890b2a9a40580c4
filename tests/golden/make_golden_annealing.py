"""
Golden vectors for the annealing caller of the time-evolution path, from the UNMODIFIED reference
(/root/reference, build container only):

    python tests/golden/make_golden_annealing.py

AnnealingProblem.compute (reference: time_evolution.py:1142-1191) ends with one static solve per problem
(static_problem.py, outside the path). The harness replaces ONLY that hand-over - TimeEvolutionProblem.
get_static_problem returns a recorder - so the stored quantities are what the reference's own loop produced:
temperature_profiles (its return value) and the vortex configuration it hands to the static solver. The stepping
arithmetic is untouched. numpy's global generator is seeded with the case seed right before compute().
"""
import os
import sys
import warnings

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness import import_reference  # noqa: E402
from tests import cases  # noqa: E402


class _Recorder:
    taken = []

    def __init__(self, n):
        _Recorder.taken.append(np.array(n))

    def compute(self):
        return None, 2, None


def main():
    ref = import_reference()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    ref.TimeEvolutionProblem.get_static_problem = \
        lambda self, vortex_configuration, problem_nr=0, time_step=0: _Recorder(vortex_configuration)
    for name, make in cases.ANNEAL_CASES.items():
        kw, seed = make(ref)
        _Recorder.taken = []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            prob = ref.AnnealingProblem(**kw)
            np.random.seed(seed)
            status, _, profiles = prob.compute()
        n = np.stack(_Recorder.taken, axis=1)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), temperature_profiles=profiles, n=n,
                            numpy_version=np.__version__, scipy_version=scipy.__version__)
        print(name, profiles.shape, n.shape, "T range", profiles.min(), profiles.max(), "vortices", np.abs(n).sum(axis=0))


if __name__ == "__main__":
    main()
