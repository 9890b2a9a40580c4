"""The numpy oracle against the golden vectors produced by the unmodified reference (and, when the
reference tree is present, against the reference itself run live)."""
import os
import warnings

import numpy as np
import pytest

import pyjjasim_b200 as pj
from oracle import oracle
from oracle.ref_harness import reference_available
from tests import cases
from tests.golden.make_golden import matrix_digest


def run_oracle(name):
    kw, seed = cases.build(name, pj)
    args, extra = cases.oracle_inputs(kw)
    W = pj.TimeEvolutionProblem(**kw).get_problem_count()
    rng = np.random.RandomState(seed) if seed is not None else None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return kw, oracle.time_evolution(*args, W, rng=rng, **extra)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_matches_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    kw, (th, I, V) = run_oracle(name)
    assert matrix_digest(kw["circuit"].get_cycle_matrix()) == str(g["A_digest"]), "cycle matrix differs from the reference's"
    for key, arr in (("theta", th), ("current", I), ("voltage", V)):
        if key in g.files:
            assert arr is not None and arr.shape == g[key].shape
            # same numpy/scipy build -> bitwise; allow round-off of a different BLAS/SuperLU build
            assert np.max(np.abs(arr - g[key])) <= 1e-11, key
        else:
            assert arr is None


def test_noise_replay_reproduces_reference_draws(golden_dir):
    # feeding the replayed draw sequence through the oracle's injection hook gives the same result as
    # drawing from the seeded generator (quirk Q2, both branches)
    for name in ("noise_small", "noise_recycled"):
        kw, seed = cases.build(name, pj)
        args, extra = cases.oracle_inputs(kw)
        prob = pj.TimeEvolutionProblem(**kw)
        W, Nt, Nj = prob.get_problem_count(), prob._Nt(), kw["circuit"]._Nj()
        Z = cases.replay_noise(Nj, W, Nt, seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            th, _, _ = oracle.time_evolution(*args, W, noise=lambda i: Z[i], **extra)
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        assert np.max(np.abs(th - g["theta"])) <= 1e-11


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_oracle_bitwise_against_live_reference():
    from oracle.ref_harness import import_reference
    ref = import_reference()
    for name in ("sq_mixed", "triangular_mutual", "noise_recycled"):
        kw_ref, seed = cases.build(name, ref)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            prob = ref.TimeEvolutionProblem(**kw_ref)
            if seed is not None:
                np.random.seed(seed)
            res = prob.compute()
        _, (th, I, V) = run_oracle(name)
        assert np.array_equal(th, res.theta)
        if res.current is not None:
            assert np.array_equal(I, res.current)
        if res.voltage is not None:
            assert np.array_equal(V, res.voltage)


def test_invariants_kcl_and_flux_quantisation():
    # KCL  M (I - Is) = 0  and flux quantisation  A (theta + L (I - Is)) + 2 pi f = 0  
    kw, _ = cases.build("sq_frustrated", pj)
    c = kw["circuit"]
    c.set_inductance(0.2)
    args, extra = cases.oracle_inputs(kw)
    extra["has_inductance"] = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, I, _ = oracle.time_evolution(*args, 6, **extra)
    Is = np.asarray(kw["current_sources"])[:, :, 0]
    M, A = c.get_cut_matrix(), c.get_cycle_matrix()
    for k in range(th.shape[2]):
        assert np.max(np.abs(M @ (I[:, :, k] - Is))) < 1e-12
        r = A @ (th[:, :, k] + c._L() @ (I[:, :, k] - Is)) + 2 * np.pi * 0.1
        assert np.max(np.abs(r)) < 1e-11


# ---------------------------------------------------------------- annealing caller (reference: time_evolution.py:1070-1191)
@pytest.mark.parametrize("name", list(cases.ANNEAL_CASES))
def test_annealing_oracle_matches_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    kw, seed = cases.ANNEAL_CASES[name](pj)
    args, extra = cases.anneal_oracle_inputs(kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prof, th, n = oracle.annealing(*args, rng=np.random.RandomState(seed), **extra)
    assert np.array_equal(prof, g["temperature_profiles"])
    assert np.array_equal(n, g["n"])
    # the temperature moved in both directions, so both branches of the adjustment rule were taken
    d = np.diff(np.vstack((np.full((1, prof.shape[1]), kw["start_T"]), prof)), axis=0)
    assert (d > 0).any() and (d < 0).any()
    # the replayed draw sequence (what the GPU parity test injects) reproduces the seeded run
    Z = cases.anneal_noise(kw["circuit"]._Nj(), kw["problem_count"], kw["interval_steps"], kw["interval_count"], seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prof2, th2, n2 = oracle.annealing(*args, noise=lambda k, s: Z[k, s], **extra)
    assert np.array_equal(prof2, prof) and np.array_equal(th2, th) and np.array_equal(n2, n)
