"""
Parity cases for the time-evolution hot path, written once against the reference's public API so the
same function builds the problem for
  * the unmodified reference (golden generation, tests/golden/make_golden.py, build container only),
  * this package (GPU parity tests),
  * the numpy oracle (CPU tests; inputs are read back from the problem object).

``build(pkg)`` returns a dict with the TimeEvolutionProblem constructor arguments; ``pkg`` is either the
reference module ``pyjjasim`` or ``pyjjasim_b200``.
"""
import numpy as np


def _sq_mixed(pkg):
    # non-uniform R, C, Ic, scalar L, per-problem f, time-dependent Is array, non-zero Vs  (SURVEY.md section 0 probe)
    a = pkg.SquareArray(9, 7)
    rng = np.random.RandomState(3)
    Nj, Nf = a._Nj(), a._Nf()
    a.set_resistance(0.5 + rng.rand(Nj))
    a.set_capacitance(0.2 + rng.rand(Nj))
    a.set_critical_current(0.6 + 0.8 * rng.rand(Nj))
    a.set_inductance(0.3)
    W, Nt = 5, 60
    f = np.linspace(0.0, 0.4, W)[None, :, None] * np.ones((Nf, 1, 1))
    Is = a.current_base(angle=0.3)[:, None, None] * (0.5 + 0.1 * np.arange(W))[None, :, None] * \
        (1.0 + 0.2 * np.sin(0.3 * np.arange(Nt)))[None, None, :]
    Vs = 0.05 * rng.randn(Nj)[:, None, None] * np.ones((1, W, 1))
    th1 = 0.1 * rng.randn(Nj, W)
    th2 = th1 + 0.01 * rng.randn(Nj, W)
    return dict(circuit=a, time_step=0.05, time_step_count=Nt, external_flux=f, current_sources=Is,
                voltage_sources=Vs, config_at_minus_1=th1, config_at_minus_2=th2)


def _sq_iv(pkg):
    # BASELINE config 1 scaled down in time: SquareArray(20,20), 32 bias currents, f=0, T=0, dt=0.05
    a = pkg.SquareArray(20, 20)
    Is = a.current_base(angle=0)[:, None, None] * np.linspace(0, 2, 32)[None, :, None]
    Nt = 400
    return dict(circuit=a, time_step=0.05, time_step_count=Nt, current_sources=Is,
                store_time_steps=[Nt // 3, Nt - 1], store_current=False, store_voltage=False)


def _sq_frustrated(pkg):
    # f != 0 (chaotic at long times; short horizon), all three outputs, sparse store mask
    a = pkg.SquareArray(12, 10)
    W, Nt = 6, 120
    Is = a.current_base(angle=0)[:, None, None] * np.linspace(0.2, 1.4, W)[None, :, None]
    return dict(circuit=a, time_step=0.05, time_step_count=Nt, external_flux=0.1, current_sources=Is,
                store_time_steps=np.arange(5, Nt, 7))


def _honeycomb(pkg):
    # BASELINE config 3 scaled down: frustrated honeycomb with DC bias (has two face-less junctions)
    a = pkg.HoneycombArray(5, 4)
    W, Nt = 8, 100
    Is = a.current_base(angle=0)[:, None, None] * np.linspace(0.1, 1.5, W)[None, :, None]
    return dict(circuit=a, time_step=0.05, time_step_count=Nt, external_flux=0.1, current_sources=Is,
                store_time_steps=np.arange(0, Nt, 10), store_voltage=False)


def _triangular_mutual(pkg):
    # general sparse inductance matrix (mutual coupling), capacitance, callable DC+AC drive (config 4 pattern)
    import scipy.sparse
    a = pkg.TriangularArray(4, 3)
    Nj = a._Nj()
    rng = np.random.RandomState(5)
    B = scipy.sparse.random(Nj, Nj, density=0.05, random_state=rng, format="csc")
    Lm = 0.02 * (B + B.T) + scipy.sparse.diags(0.5 + 0.1 * rng.rand(Nj))
    a.set_inductance(scipy.sparse.csc_matrix(Lm))
    a.set_capacitance(1.0)
    W, Nt, dt = 7, 80, 0.05
    base = a.current_base(angle=0)
    IDC = np.linspace(0, 2, W)
    Is = lambda i: base[:, None] * (IDC + 1.0 * np.sin(0.25 * i * dt))
    return dict(circuit=a, time_step=dt, time_step_count=Nt, current_sources=Is,
                store_time_steps=np.arange(0, Nt, 4))


def _noise_small(pkg):
    # T > 0, Nj <= 500: fresh draws every step (reference: time_evolution.py:536-537)
    a = pkg.SquareArray(6, 6)
    W, Nt = 4, 50
    T = np.array([0.01, 0.05, 0.1, 0.2])[None, :, None]
    return dict(circuit=a, time_step=0.05, time_step_count=Nt, external_flux=0.1, temperature=T,
                store_time_steps=np.arange(0, Nt, 5), store_voltage=False, store_current=False, _seed=11)


def _noise_recycled(pkg):
    # T > 0, Nj > 500: draws recycled with a row permutation two steps out of three (quirk Q2);
    # BASELINE config 2 pattern (f = 0.1, temperature batch, dt = 0.5) scaled down
    a = pkg.SquareArray(17, 17)
    W, Nt = 6, 30
    T = np.geomspace(1e-2, 1.0, W)[None, :, None]
    return dict(circuit=a, time_step=0.5, time_step_count=Nt, external_flux=0.1, temperature=T,
                store_time_steps=[9, 19, 29], store_voltage=False, store_current=False, _seed=7)


def _dense_sources(pkg):
    # inputs without rank-one structure: per-junction per-problem Is, per-face per-problem f, time-dependent T array
    a = pkg.SquareArray(7, 6)
    rng = np.random.RandomState(9)
    Nj, Nf = a._Nj(), a._Nf()
    W, Nt = 5, 40
    Is = 0.5 * rng.randn(Nj, W, 1)
    f = 0.2 * rng.rand(Nf, W, 1)
    return dict(circuit=a, time_step=0.05, time_step_count=Nt, external_flux=f, current_sources=Is,
                store_time_steps=np.arange(0, Nt, 3))


def _custom_cpr(pkg):
    # custom current-phase relation Ic sin^3(theta) (examples/static_example_10_custum_current_phase_relation.py:17-20)
    a = pkg.SquareArray(6, 5)
    cpr = pkg.CurrentPhaseRelation(lambda Ic, th: Ic * np.sin(th) ** 3,
                                   lambda Ic, th: 3 * Ic * np.sin(th) ** 2 * np.cos(th),
                                   lambda Ic, th: Ic * (2.0 / 3 - np.cos(th) + np.cos(th) ** 3 / 3))
    W, Nt = 4, 60
    Is = a.current_base(angle=0)[:, None, None] * np.linspace(0.3, 1.5, W)[None, :, None]
    return dict(circuit=a, time_step=0.05, time_step_count=Nt, current_phase_relation=cpr,
                external_flux=0.05, current_sources=Is, store_time_steps=np.arange(0, Nt, 6))


def _single_problem(pkg):
    # W = 1 (all inputs scalar), ragged problem count vs the 4-wide device groups
    a = pkg.SquareArray(5, 5)
    return dict(circuit=a, time_step=0.05, time_step_count=40, external_flux=0.2, current_sources=0.0,
                config_at_minus_1=0.3 * np.cos(np.arange(a._Nj()))[:, None])


CASES = {
    "sq_mixed": _sq_mixed,
    "sq_iv": _sq_iv,
    "sq_frustrated": _sq_frustrated,
    "honeycomb": _honeycomb,
    "triangular_mutual": _triangular_mutual,
    "noise_small": _noise_small,
    "noise_recycled": _noise_recycled,
    "dense_sources": _dense_sources,
    "custom_cpr": _custom_cpr,
    "single_problem": _single_problem,
}

# per-step tolerance on theta for each case: 1e-9 for non-chaotic runs, 1e-8 for the first
# few hundred steps of frustrated / noisy runs
TOL = {name: 1e-9 for name in CASES}
TOL.update(sq_frustrated=1e-8, honeycomb=1e-8, noise_small=1e-8, noise_recycled=1e-8, dense_sources=1e-8,
           sq_mixed=1e-8)


def build(name, pkg):
    kw = CASES[name](pkg)
    seed = kw.pop("_seed", None)
    return kw, seed


def replay_noise(Nj, W, Nt, seed):
    """The reference's Gaussian draw sequence (reference: time_evolution.py:533-537) as (Nt, Nj, W)."""
    rs = np.random.RandomState(seed)
    out = np.empty((Nt, Nj, W))
    rand = None
    for i in range(Nt):
        if Nj > 500:
            rand = rs.randn(Nj, W) if i % 3 == 0 else rand[rs.permutation(Nj), :]
        else:
            rand = rs.randn(Nj, W)
        out[i] = rand
    return out


def oracle_inputs(kw):
    """Constructor kwargs (built with pyjjasim_b200) -> positional/keyword arguments of oracle.time_evolution."""
    c = kw["circuit"]
    args = (c.get_cycle_matrix(), c._Ic(), c._R(), c._C(), c._L(), kw.get("time_step", 0.05),
            kw.get("time_step_count", 1000))
    cpr = kw.get("current_phase_relation", None)
    extra = dict(f=kw.get("external_flux", 0.0), Is=kw.get("current_sources", 0.0),
                 Vs=kw.get("voltage_sources", 0.0), T=kw.get("temperature", 0.0),
                 theta_m1=kw.get("config_at_minus_1", None), theta_m2=kw.get("config_at_minus_2", None),
                 store_time_steps=kw.get("store_time_steps", None), store_theta=kw.get("store_theta", True),
                 store_voltage=kw.get("store_voltage", True), store_current=kw.get("store_current", True),
                 has_inductance=c._has_inductance())
    if cpr is not None:
        extra["cpr"] = cpr.eval
    return args, extra


# ---------------------------------------------------------------------------------------------------------------
# Annealing cases (reference: time_evolution.py:1070-1191, AnnealingProblem): constructor arguments + numpy seed.
def _anneal_small(pkg):
    # Nj <= 500: fresh draws every step; the mobility target is crossed in both directions within the run
    a = pkg.SquareArray(7, 7)
    return dict(circuit=a, time_step=0.5, interval_steps=10, external_flux=0.2, current_sources=0, problem_count=6,
                interval_count=14, vortex_mobility=0.02, start_T=0.4, T_factor=1.25), 21


def _anneal_recycled(pkg):
    # Nj > 500: recycled draws (quirk Q2), per-face flux array, non-zero bias current, per-iteration mobility targets
    a = pkg.SquareArray(17, 17)
    rng = np.random.RandomState(4)
    f = 0.1 + 0.02 * rng.rand(a._Nf())
    return dict(circuit=a, time_step=0.5, interval_steps=6, external_flux=f, current_sources=0.1, problem_count=5,
                interval_count=9, vortex_mobility=np.linspace(0.03, 0.0, 9), start_T=0.3, T_factor=1.1), 22


ANNEAL_CASES = {"anneal_small": _anneal_small, "anneal_recycled": _anneal_recycled}


def anneal_noise(Nj, W, interval_steps, interval_count, seed):
    """The reference's draws during AnnealingProblem.compute as (interval_count, interval_steps, Nj, W): the global
    generator runs on across the compute() calls while the recycling pattern restarts with every call
    (reference: time_evolution.py:533-537, :1164-1174). The closing T = 0 runs draw nothing."""
    rs = np.random.RandomState(seed)
    out = np.empty((interval_count, interval_steps, Nj, W))
    for k in range(interval_count):
        rand = None
        for i in range(interval_steps):
            if Nj > 500:
                rand = rs.randn(Nj, W) if i % 3 == 0 else rand[rs.permutation(Nj), :]
            else:
                rand = rs.randn(Nj, W)
            out[k, i] = rand
    return out


def anneal_oracle_inputs(kw):
    c = kw["circuit"]
    args = (c.get_cycle_matrix(), c._Ic(), c._R(), c._C(), c._L())
    extra = {k: kw[k] for k in ("time_step", "interval_steps", "external_flux", "current_sources", "problem_count",
                                "interval_count", "start_T", "T_factor") if k in kw}
    extra["vortex_mobility_target"] = kw.get("vortex_mobility", 0.001)
    return args, extra


# ---------------------------------------------------------------------------------------------------------------
# Static polish cases (reference: static_problem.py:550-610): lattice, inductance, flux, vortex configurations, bias.
def static_cases(pkg):
    out = {}
    for name, ctor, args, L in (("square", "SquareArray", (9, 8), 0.05), ("honeycomb", "HoneycombArray", (4, 5), 0.0)):
        a = getattr(pkg, ctor)(*args)
        Nf, W = a._Nf(), 6
        rng = np.random.RandomState(0)
        f = 0.06 * np.ones((Nf, W))
        n = np.zeros((Nf, W), dtype=int)
        for w in range(1, W):
            n[rng.choice(Nf, size=w, replace=False), w] = 1
        n[rng.choice(Nf), 5] = -1
        Is = a.current_base(angle=0)[:, None] * np.linspace(0, 0.2, W)[None, :]
        out[name] = (ctor, args, L, f, n, Is)
    return out
