"""GPU parity: the device path through the public API / C ABI against the golden vectors of the unmodified
reference and against the CPU oracle, per stored step."""
import os
import warnings

import numpy as np
import pytest

import pyjjasim_b200 as pj
from oracle import oracle
from tests import cases

pytestmark = pytest.mark.gpu

ENGINES = ["streaming", "auto"]


def run_device(name, engine, **override):
    kw, seed = cases.build(name, pj)
    kw.update(override)
    if seed is not None:
        c = kw["circuit"]
        prob0 = pj.TimeEvolutionProblem(**kw)
        kw["noise_replay"] = cases.replay_noise(c._Nj(), prob0.get_problem_count(), prob0._Nt(), seed)
    os.environ["JJ_ENGINE"] = engine
    try:
        prob = pj.TimeEvolutionProblem(**kw)
        return kw, prob, prob.compute()
    finally:
        os.environ.pop("JJ_ENGINE", None)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", list(cases.CASES))
def test_device_matches_reference_golden(name, engine, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    kw, prob, res = run_device(name, engine)
    tol = cases.TOL[name]
    for key in ("theta", "current", "voltage"):
        got = getattr(res, key)
        if key in g.files:
            assert got is not None and got.shape == g[key].shape
            scale = 1.0 if key == "theta" else 1.0 / kw.get("time_step", 0.05) if key == "voltage" else 1.0
            err = np.max(np.abs(got - g[key]))
            assert err <= tol * max(scale, 1.0) * 10, (key, err)
        else:
            assert got is None
    th_err = np.max(np.abs(res.theta - g["theta"]))
    assert th_err <= tol, th_err


@pytest.mark.parametrize("engine", ENGINES)
def test_vortex_configuration_and_invariants(engine):
    kw, prob, res = run_device("sq_frustrated", engine)
    c = kw["circuit"]
    args, extra = cases.oracle_inputs(kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, I, V = oracle.time_evolution(*args, prob.get_problem_count(), **extra)
    n_dev = res.get_vortex_configuration()
    n_ora = oracle.vortex_configuration(c.get_cycle_matrix(), th)
    assert np.array_equal(n_dev, n_ora)
    Is = np.asarray(kw["current_sources"])[:, :, 0]
    M, A = c.get_cut_matrix(), c.get_cycle_matrix()
    for k in range(res.theta.shape[2]):
        assert np.max(np.abs(M @ (res.current[:, :, k] - Is))) < 1e-11
        assert np.max(np.abs(A @ res.theta[:, :, k] + 2 * np.pi * 0.1)) < 1e-10


def test_device_solve_matches_direct_solve():
    import scipy.sparse.linalg
    from pyjjasim_b200 import engine
    from pyjjasim_b200.factor import system_matrix
    a = pj.SquareArray(30, 30)
    a.set_inductance(0.1)
    tab = engine.CircuitTables(a, 0.05)
    eng = engine.DeviceEngine(0)
    eng.set_circuit(tab, pj.DefaultCPR())
    eng.set_problem(7, 0.05)
    rng = np.random.RandomState(0)
    b = rng.randn(a._Nf(), 7)
    J = eng.debug_solve(b)
    S = system_matrix(a.get_cycle_matrix(), a._L(), tab.Rv, tab.Cv)
    Jref = scipy.sparse.linalg.spsolve(S.tocsc(), b)
    eng.close()
    assert np.max(np.abs(J - Jref)) <= 1e-12 * np.max(np.abs(Jref))


@pytest.mark.parametrize("engine", ENGINES)
def test_reference_style_problem_object_through_the_device_core(engine, monkeypatch):
    # INTEGRATION.md's binding: the reference's own TimeEvolutionProblem (no _raw_sources, inputs as (N, W, Nt) broadcast
    # views or callables, time-dependence flags frozen at construction) and a circuit offering only the reference's
    # getters go through device_time_evolution_core; per-step parity with the oracle, and the fast engine must be the
    # one that ran (constants held as broadcast views are not "dense time-dependent tables")
    from pyjjasim_b200 import engine as eng_mod
    from tests.refstyle import RefStyleProblem, RefStyleCircuit
    monkeypatch.setenv("JJ_ENGINE", engine)
    a = pj.SquareArray(14, 12)
    a.set_inductance(0.05)
    W, Nt, dt = 10, 60, 0.05
    base, amps = a.current_base(angle=0), np.linspace(0.3, 1.6, W)
    Is = lambda i: base[:, None] * (amps + 0.2 * np.sin(0.4 * i * dt))[None, :]
    Vs = np.zeros((a._Nj(), 1, 1)); Vs[3] = 0.02
    kw = dict(time_step=dt, time_step_count=Nt, external_flux=0.1, current_sources=Is, voltage_sources=Vs)
    prob = RefStyleProblem(RefStyleCircuit(a), current_phase_relation=pj.DefaultCPR(), **kw)
    mask = np.zeros(Nt, dtype=bool); mask[[5, 30, Nt - 1]] = True
    th, I = eng_mod.device_time_evolution_core(prob, mask, mask)
    assert th.shape == (a._Nj(), W, 5) and I.shape == th.shape
    if engine == "auto":
        assert eng_mod.last_run_stats[0]["engine"] == 3
    args, extra = cases.oracle_inputs(dict(kw, circuit=a, store_time_steps=[5, 30, Nt - 1], store_voltage=False))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th_o, I_o, _ = oracle.time_evolution(*args, W, **extra)
    assert np.max(np.abs(th[:, :, 2:] - th_o)) <= 1e-9 and np.max(np.abs(I[:, :, 2:] - I_o)) <= 1e-9
    assert np.array_equal(th[:, :, 1], prob.config_at_minus_1) and np.array_equal(th[:, :, 0], prob.config_at_minus_2)


def test_reference_example_callable_runs_through_its_closed_form():
    # examples/time_evolution_example_5_giant_shapiro_steps.py drives the array with a plain callable
    # Is(i) = base[:, None] * (IDC + IAmp sin(f i dt)): its closed form is recovered from a few evaluations, the run
    # evaluates it a few dozen times instead of once per step, and the phases equal those of the exact rank-one input
    a = pj.SquareArray(20, 20)
    W, Nt, dt = 101, 3000, 0.05
    base, IDC = a.current_base(angle=0), np.linspace(0, 2, W)
    calls = [0]

    def Is(i):
        calls[0] += 1
        return base[:, None] * (IDC + 1.0 * np.sin(0.25 * i * dt))
    kw = dict(time_step=dt, time_step_count=Nt, store_time_steps=[Nt // 3, Nt - 1], store_current=False, store_voltage=False)
    res = pj.TimeEvolutionProblem(a, current_sources=Is, **kw).compute()
    assert calls[0] < 400                      # (constructor + classification + fit + sparse checks)
    exact = pj.TimeEvolutionProblem(a, current_sources=pj.RankOneSource(base, lambda i: IDC + 1.0 * np.sin(0.25 * i * dt)), **kw).compute()
    scale = max(1.0, float(np.max(np.abs(exact.theta))))
    assert scale > 10.0 and np.max(np.abs(res.theta - exact.theta)) <= 1e-8 * scale


def test_chunked_run_equals_single_run(monkeypatch):
    # time-dependent tables uploaded chunk by chunk give the same result as one chunk
    from pyjjasim_b200 import engine
    kw, prob, res = run_device("sq_mixed", "streaming")
    monkeypatch.setattr(engine, "_TABLE_BYTES", 5 * 8 * 7)      # 7 steps per chunk
    kw2, prob2, res2 = run_device("sq_mixed", "streaming")
    assert np.max(np.abs(res.theta - res2.theta)) <= 1e-13
    assert np.max(np.abs(res.current - res2.current)) <= 1e-13


def test_resume_from_final_state():
    # manual resume contract of the reference: pass the last two thetas as config_at_minus_1/2
    kw, seed = cases.build("sq_frustrated", pj)
    kw["store_time_steps"] = None
    kw["store_voltage"] = False
    full = pj.TimeEvolutionProblem(**kw).compute().theta
    kw1 = dict(kw, time_step_count=70)
    th1 = pj.TimeEvolutionProblem(**kw1).compute().theta
    kw2 = dict(kw, time_step_count=50, config_at_minus_1=th1[:, :, -1], config_at_minus_2=th1[:, :, -2])
    th2 = pj.TimeEvolutionProblem(**kw2).compute().theta
    assert np.max(np.abs(full[:, :, 70:] - th2)) <= 1e-12


def test_missing_gpu_library_fails_loudly(monkeypatch):
    from pyjjasim_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libjjstep.so")
    kw, _ = cases.build("single_problem", pj)
    with pytest.raises(RuntimeError):
        pj.TimeEvolutionProblem(**kw).compute()


def test_device_noise_statistics_and_shard_invariance():
    from pyjjasim_b200 import engine
    a = pj.SquareArray(12, 12)
    tab = engine.CircuitTables(a, 0.05)
    eng = engine.DeviceEngine(0)
    eng.set_circuit(tab, pj.DefaultCPR())
    eng.set_problem(64, 0.05, seed=5)
    z = np.stack([eng.debug_noise(s) for s in range(40)])
    eng.set_problem(32, 0.05, seed=5, problem_offset=32)
    z2 = eng.debug_noise(7)
    eng.close()
    assert np.array_equal(z[7][:, 32:], z2)               # keyed by the global problem index
    n = z.size
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs(np.mean(z ** 4) - 3) < 0.1 and np.max(np.abs(z)) < 7
    assert abs(np.corrcoef(z[:-1].ravel(), z[1:].ravel())[0, 1]) < 5 / np.sqrt(n)


def test_device_normals_distribution_and_tails():
    # the device draws (Philox4x32-10 + single-precision Box-Muller) as a distribution: Kolmogorov-Smirnov distance
    # from the standard normal on 1.2e7 draws, tail mass beyond 3 / 4 / 5 sigma against the exact values within their
    # Poisson noise, per-junction and per-problem means (no structure along either counter axis), and the largest
    # |z| (32-bit uniforms: the transform reaches 6.66 sigma, P(|z| > 6.66) = 2.7e-11 of the mass is cut)
    import scipy.stats
    from pyjjasim_b200 import engine
    a = pj.SquareArray(40, 40)
    tab = engine.CircuitTables(a, 0.05)
    eng = engine.DeviceEngine(0)
    eng.set_circuit(tab, pj.DefaultCPR())
    W, steps = 128, 30
    eng.set_problem(W, 0.05, seed=20261018)
    z = np.stack([eng.debug_noise(s) for s in range(steps)])           # (steps, Nj, W)
    eng.close()
    n = z.size
    assert n >= 1.1e7
    flat = np.sort(z.ravel())
    cdf = scipy.stats.norm.cdf(flat)
    grid = np.arange(1, n + 1) / n
    D = max(np.max(grid - cdf), np.max(cdf - (grid - 1.0 / n)))
    assert D < 1.63 / np.sqrt(n), D                                      # KS critical value at alpha = 0.01
    for k in (3.0, 4.0, 5.0):
        expect = 2 * scipy.stats.norm.sf(k) * n
        got = float(np.count_nonzero(np.abs(flat) > k))
        assert abs(got - expect) <= 5 * np.sqrt(expect) + 3, (k, got, expect)
    assert 4.8 < np.max(np.abs(flat)) < 6.7
    assert abs(scipy.stats.skew(flat)) < 5 * np.sqrt(6 / n) and abs(scipy.stats.kurtosis(flat)) < 5 * np.sqrt(24 / n)
    # independence across the counter axes: means per junction / per problem / per step are those of white noise
    for axis_keep, cnt in ((1, steps * W), (2, steps * a._Nj()), (0, a._Nj() * W)):
        m = z.mean(axis=tuple(i for i in range(3) if i != axis_keep))
        assert np.max(np.abs(m)) < 5.5 / np.sqrt(cnt)
        assert abs(m.var() * cnt - 1.0) < 6 * np.sqrt(2.0 / m.size)
    # neighbouring problems / junctions / steps are uncorrelated
    for u, v in ((z[:, :, :-1], z[:, :, 1:]), (z[:, :-1], z[:, 1:]), (z[:-1], z[1:])):
        assert abs(np.mean(u * v)) < 5 / np.sqrt(u.size)


# ------------------------------------------------------------------------------------------------
# subdomain engine (cooperative kernel): cut depths / chunk widths against a direct solve and the golden vectors
# ------------------------------------------------------------------------------------------------
SUBDOMAIN_CONFIGS = ["0,1", "1,2", "2,4", "3,4", "3,8", "4,1"]


@pytest.mark.parametrize("W", [13, 70])
@pytest.mark.parametrize("cfg", SUBDOMAIN_CONFIGS)
def test_subdomain_solve_matches_direct_solve(cfg, W):
    _subdomain_solve_check(cfg, W)


@pytest.mark.parametrize("tt_max", ["0", "40", "150"])
@pytest.mark.parametrize("cfg,W", [("3,4", 70), ("4,1", 13), ("4,2", 40), ("5,4", 33), ("p24,4", 100), ("p7,2", 64), ("5,8", 64)])
def test_subdomain_solve_through_upper_phases(cfg, W, tt_max, monkeypatch):
    # the separators between the subdomains and the dense top of the top are swept by the upper program (gathered
    # dense products, two phases per tree depth); tt_max = 0: no dense part at all
    monkeypatch.setenv("JJ_TT_MAX", tt_max)
    _subdomain_solve_check(cfg, W, upper=True)


def _subdomain_solve_check(cfg, W, upper=False):
    import scipy.sparse.linalg
    from pyjjasim_b200 import engine
    from pyjjasim_b200.factor import system_matrix
    d, NG = cfg.split(",")
    n_parts = int(d[1:]) if d.startswith("p") else None
    d, NG = (None if n_parts else int(d)), int(NG)
    a = pj.SquareArray(40, 37) if not upper else pj.SquareArray(61, 53)
    rng = np.random.RandomState(2)
    a.set_resistance(0.5 + rng.rand(a._Nj()))
    a.set_inductance(0.1)
    tab = engine.CircuitTables(a, 0.05, n_parts=n_parts)
    eng = engine.DeviceEngine(0)
    eng.set_circuit(tab, pj.DefaultCPR(), with_program=False)
    eng.set_subdomain(d, NG)
    plan = tab.subdomain_plan(d, NG)
    if upper:
        assert plan.upper["n_fwd"] > 0 and plan.upper["n_bwd"] > 0 and plan.tt0 > 0
    eng.set_problem(W, 0.05)
    b = rng.randn(a._Nf(), W)
    J = eng.debug_subdomain_solve(b)
    S = system_matrix(a.get_cycle_matrix(), a._L(), tab.Rv, tab.Cv)
    Jref = scipy.sparse.linalg.spsolve(S.tocsc(), b)
    eng.close()
    assert np.max(np.abs(J - Jref)) <= 1e-12 * np.max(np.abs(Jref))


@pytest.mark.parametrize("cfg", ["0,1", "1,2", "2,4", "3,1", "3,4,tt0", "4,2,tt12"])
@pytest.mark.parametrize("name", ["sq_mixed", "sq_frustrated", "honeycomb", "noise_recycled", "custom_cpr", "sq_iv"])
def test_subdomain_engine_matches_reference_golden(name, cfg, golden_dir, monkeypatch):
    if "tt" in cfg:          # run the upper program inside the time loop (tt0: no dense top at all)
        cfg, tt = cfg.rsplit(",tt", 1)
        monkeypatch.setenv("JJ_TT_MAX", tt)
    monkeypatch.setenv("JJ_SUBDOMAIN", cfg)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    kw, prob, res = run_device(name, "subdomain")
    from pyjjasim_b200 import engine
    st = engine.last_run_stats[0]
    d, NG = (int(v) for v in cfg.split(","))
    assert st["engine"] == 3 and (st["cluster_size"], st["tile_problems"]) == (1 << d, 8 * NG)
    tol = cases.TOL[name]
    assert np.max(np.abs(res.theta - g["theta"])) <= tol
    if "current" in g.files:
        assert np.max(np.abs(res.current - g["current"])) <= 10 * tol
    if "voltage" in g.files:
        assert np.max(np.abs(res.voltage - g["voltage"])) <= 10 * tol / kw.get("time_step", 0.05)


def test_subdomain_multiple_items_per_block(monkeypatch):
    # more (subdomain, chunk) items than thread blocks: z goes through global memory; same trajectories
    monkeypatch.setenv("JJ_SUBDOMAIN", "2,1")
    kw, _ = cases.build("noise_small", pj)
    kw["noise_seed"] = 7
    out = {}
    for grid in ("0", "3"):
        monkeypatch.setenv("JJ_SUB_GRID", grid)
        monkeypatch.setenv("JJ_ENGINE", "subdomain")
        out[grid] = pj.TimeEvolutionProblem(**kw).compute().theta
    assert np.array_equal(out["0"], out["3"])


def test_subdomain_and_streaming_agree_on_device_noise(monkeypatch):
    kw, _ = cases.build("noise_small", pj)
    kw["noise_seed"] = 99
    out = {}
    for eng_name in ("streaming", "subdomain"):
        monkeypatch.setenv("JJ_ENGINE", eng_name)
        out[eng_name] = pj.TimeEvolutionProblem(**kw).compute().theta
    assert np.max(np.abs(out["streaming"] - out["subdomain"])) <= 1e-9


# ---------------------------------------------------------------- annealing caller (reference: time_evolution.py:1070-1191)
def _anneal(name, engine, **extra_kw):
    kw, seed = cases.ANNEAL_CASES[name](pj)
    Z = cases.anneal_noise(kw["circuit"]._Nj(), kw["problem_count"], kw["interval_steps"], kw["interval_count"], seed)
    os.environ["JJ_ENGINE"] = engine
    try:
        ap = pj.AnnealingProblem(noise_replay=Z, **kw, **extra_kw)
        return kw, Z, ap, ap.anneal()
    finally:
        os.environ.pop("JJ_ENGINE", None)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", list(cases.ANNEAL_CASES))
def test_annealing_matches_reference_golden(name, engine, golden_dir):
    # the reference's own draws injected: the temperature schedule (decided by exact integer vortex-mobility sums)
    # and the final vortex configuration equal the unmodified reference's; phases agree with the oracle
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    kw, Z, ap, (theta, n, profiles) = _anneal(name, engine)
    assert np.array_equal(profiles, g["temperature_profiles"])
    assert np.array_equal(n, g["n"])
    args, extra = cases.anneal_oracle_inputs(kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prof_o, th_o, n_o = oracle.annealing(*args, noise=lambda k, s: Z[k, s], **extra)
    assert np.max(np.abs(theta - th_o)) <= 1e-7          # ~150 noisy steps; tolerance stated in SURVEY.md section 8c
    assert np.array_equal(ap.T[0, :, 0], profiles[-1])
    # compute(): the anneal, then the reference's closing step - one stationary-state solve per annealed configuration
    status, configs, prof2 = pj.AnnealingProblem(noise_replay=Z, **kw).compute()
    assert np.array_equal(prof2, profiles) and len(configs) == kw["problem_count"] and set(status) <= {0, 1, 2}
    assert np.array_equal(configs[1].get_vortex_configuration(), n[:, 1])
    A, M = kw["circuit"].get_cycle_matrix(), kw["circuit"].get_cut_matrix()
    assert np.all(status == 0) if name == "anneal_small" else True       # (the short, hot anneal_recycled run leaves states the Newton iteration leaves: status 1)
    for p in np.flatnonzero(status == 0):
        th_p = configs[p].get_theta()                      # a stationary state carrying the annealed vortices
        assert np.array_equal(-(A @ np.round(th_p / (2 * np.pi))).astype(int), n[:, p])
        assert np.max(np.abs(M @ (configs[p].get_current() - np.broadcast_to(configs[p].current_sources, th_p.shape)))) < 1e-8
        assert np.max(np.abs(configs[p].annealed_theta - theta[:, p])) <= 1e-7      # (this run used the default engine)
    st2, cf2, _ = pj.AnnealingProblem(noise_replay=Z, **kw).compute(polish=False)
    assert np.all(st2 == 2) and np.max(np.abs(cf2[0].get_theta() - theta[:, 0])) <= 1e-7


@pytest.mark.parametrize("engine", ENGINES)
def test_device_side_temperature_schedule_equals_the_host_rule_bit_for_bit(engine, monkeypatch):
    # jj_anneal keeps amplitudes, restarts, mobility sums and the temperature rule on the device; the same schedule with
    # one host round trip per interval (the reference's expressions in numpy on the same integer sums) must give the
    # same temperature profiles and phases exactly - both directions of the rule, per-interval targets, a cold start
    monkeypatch.setenv("JJ_ENGINE", engine)
    a = pj.SquareArray(14, 13)
    for kw in (dict(vortex_mobility=0.02, start_T=0.4, T_factor=1.25, interval_count=16),
               dict(vortex_mobility=np.linspace(0.05, 0.001, 9), start_T=0.15, T_factor=1.5, interval_count=9),
               dict(vortex_mobility=0.01, start_T=0.0, T_factor=1.1, interval_count=3)):
        args = dict(circuit=a, time_step=0.5, interval_steps=10, external_flux=0.2, current_sources=0, problem_count=10,
                    noise_seed=99, **kw)
        monkeypatch.setenv("JJ_ANNEAL_HOST", "1")
        th_h, n_h, prof_h = pj.AnnealingProblem(**args).anneal()
        monkeypatch.setenv("JJ_ANNEAL_HOST", "0")
        ap = pj.AnnealingProblem(**args)
        th_d, n_d, prof_d = ap.anneal()
        assert np.array_equal(prof_d, prof_h) and np.array_equal(th_d, th_h) and np.array_equal(n_d, n_h)
        assert np.array_equal(ap.T[0, :, 0], prof_h[-1])
        if kw["start_T"] > 0:
            ratio = prof_h[1:] / prof_h[:-1]
            assert np.any(ratio > 1) and (np.any(ratio < 1) or kw["interval_count"] < 10)


@pytest.mark.parametrize("cfg", [
    dict(array=(14, 13), W=10, env={}),                                            # one block per item, few subdomains
    dict(array=(40, 37), W=40, env={"JJ_SUBDOMAIN": "3,1"}),                       # lean kernel, 8 subdomains, chunk-local barriers
    dict(array=(40, 37), W=70, env={"JJ_SUBDOMAIN": "3,2", "JJ_SUB_GRID": "5"}),   # several items per block (work counter)
    dict(array=(61, 53), W=33, env={"JJ_SUBDOMAIN": "5,4", "JJ_TT_MAX": "40"}),    # upper phases inside the time loop
    dict(array=(61, 53), W=20, env={"JJ_SUBDOMAIN": "4,1", "JJ_TT_MAX": "0"}),     # no dense top at all
    dict(array=(14, 13), W=12, env={}, Is="drive"),                                 # driven far above Ic: zones wrap past 127
])
def test_mobility_from_phase_zone_planes_equals_the_mobility_from_phase_planes(cfg, monkeypatch):
    # on the subdomain engine jj_anneal lets the step kernel store the phase zones round(theta / 2 pi) of the interval's
    # steps as bytes and k_zone_mobility works on them modulo 256; with JJ_ANNEAL_ZONES=0 it stores the phases themselves
    # and k_vortex_mobility reads them back. Integer sums: the temperature profiles and the phases must be identical.
    from pyjjasim_b200 import engine as eng_mod
    monkeypatch.setenv("JJ_ENGINE", "subdomain")
    for k, v in cfg["env"].items():
        monkeypatch.setenv(k, v)
    a = pj.SquareArray(*cfg["array"])
    if cfg.get("Is") == "drive":
        cfg = dict(cfg, Is=30.0 * a.current_base(angle=0)[:, None, None])        # ~30 rad per unit time: > 128 zones within the schedule
    args = dict(circuit=a, time_step=0.5, interval_steps=7, external_flux=0.2, current_sources=cfg.get("Is", 0), problem_count=cfg["W"],
                noise_seed=5, vortex_mobility=0.02, start_T=0.4, T_factor=1.25, interval_count=12)
    out = {}
    for zones in ("0", "1"):
        monkeypatch.setenv("JJ_ANNEAL_ZONES", zones)
        out[zones] = pj.AnnealingProblem(**args).anneal()
        assert eng_mod.last_run_stats[0]["engine"] == 3
    (th_p, n_p, prof_p), (th_k, n_k, prof_k) = out["0"], out["1"]
    ratio = prof_p[1:] / prof_p[:-1]
    if np.ndim(cfg.get("Is", 0)) == 0:
        assert np.any(ratio > 1) and np.any(ratio < 1)            # both directions of the rule were exercised
    else:
        assert np.abs(th_p).max() > 2 * np.pi * 200               # the driven junctions wound through > 128 zones
    assert np.array_equal(prof_k, prof_p) and np.array_equal(th_k, th_p) and np.array_equal(n_k, n_p)


@pytest.mark.parametrize("engine", ENGINES)
def test_mobility_kernels_on_faces_longer_than_their_fast_path(engine, monkeypatch):
    # two faces of 11 junctions each (a ring of 20 with one chord): the mobility kernels keep up to 8 junction rows of a face
    # in registers and take a general loop beyond; the schedule must equal the host rule on the same integer sums, with
    # phase-zone bytes (subdomain engine) and with phase planes
    from pyjjasim_b200 import engine as eng_mod
    monkeypatch.setenv("JJ_ENGINE", engine)
    N = 20
    ang = 2 * np.pi * np.arange(N) / N
    n1 = np.concatenate([np.arange(N), [0]])
    n2 = np.concatenate([(np.arange(N) + 1) % N, [N // 2]])
    c = pj.Circuit(pj.EmbeddedGraph(np.cos(ang), np.sin(ang), n1, n2))
    assert np.all(np.diff(c.get_cycle_matrix().tocsr().indptr) == 11)
    args = dict(circuit=c, time_step=0.5, interval_steps=9, external_flux=0.4, current_sources=0, problem_count=7,
                noise_seed=11, vortex_mobility=0.05, start_T=1.5, T_factor=1.2, interval_count=14)
    monkeypatch.setenv("JJ_ANNEAL_HOST", "1")
    th_h, n_h, prof_h = pj.AnnealingProblem(**args).anneal()
    monkeypatch.setenv("JJ_ANNEAL_HOST", "0")
    for zones in ("1", "0"):
        monkeypatch.setenv("JJ_ANNEAL_ZONES", zones)
        th_d, n_d, prof_d = pj.AnnealingProblem(**args).anneal()
        assert np.array_equal(prof_d, prof_h) and np.array_equal(th_d, th_h) and np.array_equal(n_d, n_h)
        assert eng_mod.last_run_stats[0]["engine"] == (3 if engine == "auto" else 1)      # (zone bytes: subdomain engine only)
    ratio = prof_h[1:] / prof_h[:-1]
    assert np.any(ratio < 1)                                  # the hot start did move vortices
    # and k_vortex_mobility itself against numpy on host copies of the phases: the same draws replayed through the
    # device loop and through the loop as the reference writes it (compute() per interval, mobility on the host)
    Z = np.random.RandomState(3).randn(args["interval_count"], args["interval_steps"], c.junction_count(), args["problem_count"])
    kw = {k: v for k, v in args.items() if k != "noise_seed"}
    _, _, prof_r = pj.AnnealingProblem(noise_replay=Z, **kw).anneal()
    ref_style = pj.AnnealingProblem(**kw)
    th = np.zeros((c.junction_count(), kw["problem_count"]))
    prob = pj.TimeEvolutionProblem(c, time_step_count=kw["interval_steps"], time_step=kw["time_step"],
                                   external_flux=np.atleast_1d(kw["external_flux"])[:, None, None], current_sources=0,
                                   temperature=ref_style.T, store_current=False, store_voltage=False, stencil_width=3)
    prof = np.zeros_like(prof_r)
    for i in range(kw["interval_count"]):
        prob.temperature = ref_style.T * np.ones((1, 1, kw["interval_steps"]))
        prob.config_at_minus_1, prob.config_at_minus_2, prob.noise_replay = th, th.copy(), Z[i]
        out = prob.compute()
        ref_style._temperature_adjustment(ref_style.get_vortex_mobility(out.get_vortex_configuration()), i)
        th = out.get_theta()[..., -1]
        prof[i, :] = ref_style.T[0, :, 0]
    assert np.array_equal(prof, prof_r)


def test_phase_zones_at_the_rounding_ties():
    # the device takes round(theta / 2 pi) from a reciprocal product when that is safely away from a tie and from the
    # true division otherwise: phases placed on and next to the half-integers must give numpy's integers
    from pyjjasim_b200 import engine as eng_mod
    a = pj.SquareArray(3, 3)
    Nj, W = a.junction_count(), 64
    rng = np.random.RandomState(0)
    k = rng.randint(-2000, 2000, size=(Nj, W)).astype(np.double)
    th = (k + 0.5) * (2 * np.pi)
    for _ in range(3):                                   # a few ulps to either side of the tie, and the tie itself
        step = rng.randint(-3, 4, size=th.shape)
        th = np.where(step > 0, np.nextafter(th, np.inf), np.where(step < 0, np.nextafter(th, -np.inf), th))
    th[:, ::7] = rng.randn(Nj, th[:, ::7].shape[1]) * 1e7      # huge phases: always the division
    tab = eng_mod._tables_for(a, 0.05)
    key, e, kind = eng_mod._engine_for(tab, pj.DefaultCPR(), 0, W, eng_mod._engine_kind(None))
    try:
        e.set_problem(W, 0.05, engine=kind)
        e.set_state(th, th)
        want = -(a.get_cycle_matrix() @ np.round(th / (2.0 * np.pi))).astype(int)
        assert np.array_equal(e.vortex_configuration(-1), want)
    finally:
        eng_mod._release_engine(0, key, e, True)


def test_annealing_equals_reference_style_loop():
    # the loop exactly as the reference writes it (re-entering compute(), assigning prob.temperature and the
    # initial conditions, vortex configurations on the host) gives the same schedule as the device-resident path
    kw, Z, ap, (theta, n, profiles) = _anneal("anneal_small", "auto")
    ref_style = pj.AnnealingProblem(**kw)
    f = np.atleast_1d(kw["external_flux"])[:, None, None]
    th = np.zeros((kw["circuit"].junction_count(), kw["problem_count"]))
    prob = pj.TimeEvolutionProblem(kw["circuit"], time_step_count=kw["interval_steps"], time_step=kw["time_step"],
                                   external_flux=f, current_sources=kw["current_sources"], temperature=ref_style.T,
                                   store_current=False, store_voltage=False, stencil_width=3)
    prof = np.zeros_like(profiles)
    for i in range(kw["interval_count"]):
        prob.temperature = ref_style.T * np.ones((1, 1, kw["interval_steps"]))
        prob.config_at_minus_1 = th
        prob.config_at_minus_2 = th.copy()
        prob.noise_replay = Z[i]
        out = prob.compute()
        ref_style._temperature_adjustment(ref_style.get_vortex_mobility(out.get_vortex_configuration()), i)
        th = out.get_theta()[..., -1]
        prof[i, :] = ref_style.T[0, :, 0]
    assert np.array_equal(prof, profiles)


@pytest.mark.parametrize("engine", ENGINES)
def test_device_vortex_observables_are_exact(engine):
    # n = -A round(theta / 2 pi) and the mobility sums computed on the device from stored theta planes equal the
    # reference's formulas applied on the host to the very same planes (integers: exact)
    from pyjjasim_b200 import engine as eng_mod
    kw, _ = cases.build("sq_frustrated", pj)
    c, W, Nt = kw["circuit"], 6, 60
    kw.update(store_time_steps=None, store_current=False, store_voltage=False, time_step_count=Nt,
              current_sources=c.current_base(angle=0)[:, None, None] * np.linspace(0.9, 2.4, W)[None, :, None])
    prob = pj.TimeEvolutionProblem(**kw)
    os.environ["JJ_ENGINE"] = engine
    try:
        tab = eng_mod._tables_for(c, 0.05)
        key, e, kind = eng_mod._engine_for(tab, pj.DefaultCPR(), 0, W, eng_mod._engine_kind(None))
        e.set_problem(W, 0.05, engine=kind)
        e.alloc_outputs(Nt, 0)
        specs = eng_mod._classify_all(prob, tab)
        eng_mod._setup_sources(e, specs, eng_mod._ShardInputs(specs, 0, W), tab)
        e.run(0, Nt, np.arange(Nt), None)
        theta = np.moveaxis(e.fetch_theta(0, Nt), 0, 2)
        n_host = oracle.vortex_configuration(c.get_cycle_matrix(), theta)
        assert np.abs(np.diff(n_host, axis=2)).sum() > 0        # vortices do move in this run
        for k in (0, Nt // 2, Nt - 1):
            assert np.array_equal(e.vortex_configuration(k), n_host[:, :, k])
        assert np.array_equal(e.vortex_configuration(-1), n_host[:, :, -1])
        assert np.array_equal(e.vortex_mobility_sums(0, Nt), np.abs(np.diff(n_host, axis=2)).sum(axis=(0, 2)))
        assert np.array_equal(e.vortex_mobility_sums(3, 10), np.abs(np.diff(n_host[:, :, 3:13], axis=2)).sum(axis=(0, 2)))
        assert np.array_equal(e.vortex_mobility_sums(5, 1), np.zeros(W, dtype=int))
        a1, a2 = e.get_state()
        e.restart_at_rest()
        b1, b2 = e.get_state()
        assert np.array_equal(b1, a1) and np.array_equal(b2, a1) and not np.array_equal(a2, a1)
        eng_mod._release_engine(0, key, e, True)
    finally:
        os.environ.pop("JJ_ENGINE", None)


@pytest.mark.parametrize("tt_max", [None, "120", "0"])
def test_larger_circuit_runs_on_subdomain_engine_with_several_items_per_block(tt_max, monkeypatch):
    # a circuit too large for one subdomain per (SM, chunk) pair (the cfg3 / cfg4 regime, scaled down): the layout cuts
    # it finer, every block loops over several (subdomain, chunk) items per time step; with a small dense top of the top
    # (tt_max) the separators above the subdomains go through the upper program (block and warp tasks) every time step;
    # per-step parity with the oracle
    from pyjjasim_b200 import engine
    if tt_max is not None:
        monkeypatch.setenv("JJ_TT_MAX", tt_max)
    a = pj.SquareArray(80, 80)
    W, Nt = 512, 12
    NG, chunks, n_parts = engine.subdomain_layout(a._Nf(), W, engine._sm_count(0))
    assert n_parts * chunks > engine._sm_count(0)
    Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0.2, 1.8, W))
    kw = dict(circuit=a, time_step=0.05, time_step_count=Nt, external_flux=0.1, current_sources=Is,
              store_time_steps=[3, Nt - 1], store_current=False, store_voltage=False)
    prob = pj.TimeEvolutionProblem(**kw)
    res = prob.compute()
    st = engine.last_run_stats[0]
    assert st["engine"] == 3 and st["cluster_size"] == n_parts
    if tt_max is not None:
        up = engine._tables_for(a, 0.05, n_parts).subdomain_plan(None, NG, chunks).upper
        assert up["n_fwd"] > 0 and up["n_bwd"] > 0
    kw["current_sources"] = (a.current_base(angle=0)[:, None] * np.linspace(0.2, 1.8, W)[None, :])[:, :, None]
    args, extra = cases.oracle_inputs(kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, _, _ = oracle.time_evolution(*args, W, **extra)
    assert np.max(np.abs(res.theta - th)) <= 1e-9


# ---------------------------------------------------------------- full-size configurations: size-independent properties
def _flux_and_kcl(a, res, f, Is_last):
    # flux quantisation A theta + 2 pi f = 0 and Kirchhoff's current law M (I - Is) = 0 hold at every stored step
    # whatever happened before it: a solve that is wrong anywhere in the circuit breaks them
    A, M = a.get_cycle_matrix(), a.get_cut_matrix()
    th, I = res.theta[:, :, -1], res.current[:, :, -1]
    scale = max(1.0, float(np.max(np.abs(th))))
    return (float(np.max(np.abs(A @ th + 2 * np.pi * f))) / scale, float(np.max(np.abs(M @ (I - Is_last)))))


def test_full_size_cfg2_invariants_and_chunking():
    # BASELINE config 2 at full size (SquareArray(100,100), 256 temperatures, f = 0.1, dt = 0.5) over a short horizon
    a = pj.SquareArray(100, 100)
    W, Nt = 256, 60
    T = np.geomspace(1e-2, 1.0, W)[None, :, None]
    kw = dict(circuit=a, time_step=0.5, time_step_count=Nt, external_flux=0.1, temperature=T, noise_seed=1234,
              store_time_steps=[Nt // 3, Nt - 1], store_voltage=False)
    res = pj.TimeEvolutionProblem(**kw).compute()
    from pyjjasim_b200 import engine
    assert engine.last_run_stats[0]["engine"] == 3
    flux, kcl = _flux_and_kcl(a, res, 0.1, 0.0)
    assert flux < 1e-10 and kcl < 1e-9
    # the hotter problems hold more vortices at the end than the coldest ones (the noise really is per problem)
    n = res.get_vortex_configuration()[:, :, -1]
    assert np.abs(n[:, -32:]).sum() > np.abs(n[:, :32]).sum()
    assert np.all(np.isfinite(res.theta))


def test_full_size_cfg4_share_invariants():
    # BASELINE config 4 (SquareArray(256,256) with capacitance, DC + AC drive) with one GPU's share of 512 problems:
    # 128 subdomains, ~14 items per block, ~5 500 separator rows above them (upper program + dense top of the top);
    # a few steps, then the invariants
    a = pj.SquareArray(256, 256)
    a.set_capacitance(1.0)
    W, Nt, dt = 512, 6, 0.05
    base = a.current_base(angle=0)
    IDC, IA = np.linspace(0, 2, W), np.linspace(0, 3, W)
    Is = pj.RankOneSource(base, lambda i: IDC + IA * np.sin(0.25 * i * dt), problem_count=W)
    res = pj.TimeEvolutionProblem(a, time_step=dt, time_step_count=Nt, current_sources=Is, external_flux=0.05,
                                  temperature=0.01 * np.ones((1, W, 1)), noise_seed=5, store_time_steps=[Nt - 1],
                                  store_voltage=False).compute()
    from pyjjasim_b200 import engine
    st = engine.last_run_stats[0]
    assert st["engine"] == 3 and st["cluster_size"] == engine.subdomain_layout(a._Nf(), W, engine._sm_count(0))[2]
    flux, kcl = _flux_and_kcl(a, res, 0.05, Is(Nt - 1))
    assert flux < 1e-10 and kcl < 1e-9


# ---------------------------------------------------------------- running observables (north star (3))
def _observer_case(a, W, Nt, **extra):
    base = a.current_base(angle=0)
    amps = np.linspace(0.4, 2.2, W)                       # from pinned to running problems: vortices move, phases wind
    return dict(circuit=a, time_step=0.1, time_step_count=Nt, external_flux=0.15,
                current_sources=pj.RankOneSource(base, amps), store_current=False, store_voltage=False, **extra), base, amps


def _check_observers(res, a, theta_all, first, k, Nt, dt):
    steps = np.arange(first, Nt, k)
    assert res.get_observation_count() == steps.size and np.array_equal(res.get_observed_steps(), steps)
    n_all = oracle.vortex_configuration(a.get_cycle_matrix(), theta_all[:, :, steps])
    assert np.abs(n_all).sum() > 0 and np.abs(np.diff(n_all, axis=2)).sum() > 0       # vortices are there and move
    assert np.array_equal(res.get_vortex_sum(), n_all.sum(axis=2))
    assert np.allclose(res.get_mean_vortex_configuration(), n_all.mean(axis=2), rtol=0, atol=1e-15)
    span = (steps[-1] - steps[0]) * dt
    ob = res.observed
    assert np.array_equal(ob["theta_first"], theta_all[:, :, steps[0]]), np.argwhere(ob["theta_first"] != theta_all[:, :, steps[0]])[:8]
    assert np.array_equal(ob["theta_latest"], theta_all[:, :, steps[-1]]), np.argwhere(ob["theta_latest"] != theta_all[:, :, steps[-1]])[:8]
    assert np.array_equal(res.get_dc_voltage(), (theta_all[:, :, steps[-1]] - theta_all[:, :, steps[0]]) / span)


@pytest.mark.parametrize("engine", ENGINES)
def test_running_observables_equal_what_the_stored_planes_give(engine, monkeypatch):
    # vortex sums and phase marks accumulated inside the step kernel (no plane stored) against the same quantities
    # formed on the host from ALL phase planes of the same run (exact: integers and the same two phases), including
    # ragged problem counts, an offset first observation, and chunked jj_run calls
    from pyjjasim_b200 import engine as eng_mod
    monkeypatch.setenv("JJ_ENGINE", engine)
    a = pj.SquareArray(13, 11)
    W, Nt, first, k = 10, 90, 7, 4
    kw, base, amps = _observer_case(a, W, Nt)
    full = pj.TimeEvolutionProblem(**kw).compute()
    res = pj.TimeEvolutionProblem(observe_interval=k, observe_first=first, **dict(kw, store_theta=False)).compute()
    assert res.theta is None
    _check_observers(res, a, full.theta, first, k, Nt, 0.1)
    # together with stored planes and time-dependent tables uploaded in chunks of 7 steps
    monkeypatch.setattr(eng_mod, "_TABLE_BYTES", W * 8 * 7)
    f_t = 0.15 + 0.01 * np.sin(0.3 * np.arange(Nt))[None, None, :] * np.ones((1, W, 1))
    kw2 = dict(kw, external_flux=f_t, store_time_steps=[5, 50])
    full2 = pj.TimeEvolutionProblem(**dict(kw2, store_time_steps=None)).compute()
    res2 = pj.TimeEvolutionProblem(observe_interval=k, observe_first=first, **kw2).compute()
    assert np.array_equal(res2.theta, full2.theta[:, :, [5, 50]])
    _check_observers(res2, a, full2.theta, first, k, Nt, 0.1)
    # against the oracle (non-chaotic over this horizon): same integers
    args, extra = cases.oracle_inputs(dict(kw, current_sources=(base[:, None] * amps[None, :])[:, :, None]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, _, _ = oracle.time_evolution(*args, W, **extra)
    steps = np.arange(first, Nt, k)
    assert np.array_equal(res.get_vortex_sum(), oracle.vortex_configuration(a.get_cycle_matrix(), th[:, :, steps]).sum(axis=2))
    assert np.max(np.abs(res.get_dc_voltage() - (th[:, :, steps[-1]] - th[:, :, steps[0]]) / ((steps[-1] - steps[0]) * 0.1))) < 1e-9


def test_running_observables_with_several_items_per_block_and_upper_phases(monkeypatch):
    # the cfg3 / cfg4 / cfg5 regime scaled down: several (subdomain, chunk) items per block, halo rows shared by
    # subdomains (integer atomics), observation at the last step of the run
    from pyjjasim_b200 import engine as eng_mod
    monkeypatch.setenv("JJ_TT_MAX", "120")
    a = pj.SquareArray(80, 80)
    W, Nt, first, k = 512, 13, 0, 3
    kw, base, amps = _observer_case(a, W, Nt)
    full = pj.TimeEvolutionProblem(**kw).compute()
    assert eng_mod.last_run_stats[0]["engine"] == 3
    res = pj.TimeEvolutionProblem(observe_interval=k, observe_first=first, **dict(kw, store_theta=False)).compute()
    _check_observers(res, a, full.theta, first, k, Nt, 0.1)


@pytest.mark.parametrize("engine", ENGINES)
def test_vortex_configurations_of_the_stored_steps_without_the_phases_leaving_the_device(engine, monkeypatch):
    from pyjjasim_b200 import engine as eng_mod
    monkeypatch.setenv("JJ_ENGINE", engine)
    a = pj.HoneycombArray(6, 7)
    W, Nt = 9, 60
    kw, base, amps = _observer_case(a, W, Nt, store_time_steps=[0, 11, 12, 40, 59])
    want = pj.TimeEvolutionProblem(**kw).compute().get_vortex_configuration()
    assert np.abs(want).sum() > 0
    res = pj.TimeEvolutionProblem(store_vortex_configuration=True, **dict(kw, store_theta=False)).compute()
    assert res.theta is None and np.array_equal(res.get_vortex_configuration(), want)
    assert np.array_equal(res.get_vortex_configuration([12, 59]), want[:, :, [2, 4]])
    with pytest.raises(pj.ThetaNotStored):
        res.get_theta()
    # the planes one jj_run call keeps on the device are bounded: a tiny budget forces one call per stored plane
    monkeypatch.setattr(eng_mod, "_PLANE_BYTES", a._Nj() * W * 8)
    res2 = pj.TimeEvolutionProblem(store_vortex_configuration=True, **kw).compute()
    assert np.array_equal(res2.get_vortex_configuration(), want)
    assert np.array_equal(res2.theta, pj.TimeEvolutionProblem(**kw).compute().theta)


def _oracle_sub_batch(a, kw, base, amps, sel, **oracle_extra):
    """oracle.time_evolution on the problems `sel` of a rank-one current sweep base[e] * amps[w] (the oracle's cost is
    per problem: a sub-batch of a full-size device run is compared, the device run itself uses the full batch and
    therefore the plan, chunking and item schedule of the named configuration)"""
    kw = dict(kw)
    kw["current_sources"] = (base[:, None] * amps[sel][None, :])[:, :, None]
    args, extra = cases.oracle_inputs(kw)
    extra.update(oracle_extra)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, _, _ = oracle.time_evolution(*args, len(sel), **extra)
    return th


def test_full_size_cfg3_per_step_against_oracle():
    # BASELINE config 3 at full size: HoneycombArray(200,200) (239 400 junctions), f = 0.1, DC bias sweep, 512 problems:
    # 256 subdomains, upper program + dense top of the top. Per-step parity with the oracle on problems from both ends
    # and the middle of the sweep, over a short horizon (frustrated: 1e-8)
    from pyjjasim_b200 import engine
    a = pj.HoneycombArray(200, 200)
    W, Nt, dt = 512, 12, 0.05
    base, amps = a.current_base(angle=0), np.linspace(0.1, 1.5, W)
    kw = dict(circuit=a, time_step=dt, time_step_count=Nt, external_flux=0.1, store_time_steps=[4, Nt - 1],
              store_current=False, store_voltage=False)
    res = pj.TimeEvolutionProblem(current_sources=pj.RankOneSource(base, amps), **kw).compute()
    st = engine.last_run_stats[0]
    assert st["engine"] == 3 and st["cluster_size"] == engine.subdomain_layout(a._Nf(), W, engine._sm_count(0))[2]
    sel = np.array([0, 1, 255, 256, 257, 509, 510, 511])
    th = _oracle_sub_batch(a, kw, base, amps, sel)
    assert np.max(np.abs(th)) > 0.5
    assert np.max(np.abs(res.theta[:, sel, :] - th)) <= 1e-8


def test_full_size_cfg4_per_step_against_oracle():
    # BASELINE config 4: SquareArray(256,256) with capacitance, DC + AC drive, one GPU's share of 512 problems
    # (T = 0 here: trajectories with device noise cannot be compared step by step; the noisy variant is covered by the
    # invariants above)
    from pyjjasim_b200 import engine
    a = pj.SquareArray(256, 256)
    a.set_capacitance(1.0)
    W, Nt, dt = 512, 16, 0.05
    base = a.current_base(angle=0)
    IDC, IA = np.linspace(0, 2, W), np.linspace(0, 3, W)
    Is = pj.RankOneSource(base, lambda i: IDC + IA * np.sin(0.25 * i * dt), problem_count=W)
    kw = dict(circuit=a, time_step=dt, time_step_count=Nt, external_flux=0.05, store_time_steps=[5, Nt - 1],
              store_current=False, store_voltage=False)
    res = pj.TimeEvolutionProblem(current_sources=Is, **kw).compute()
    assert engine.last_run_stats[0]["engine"] == 3
    sel = np.array([0, 100, 255, 256, 400, 511])
    kw_o = dict(kw, current_sources=lambda i: base[:, None] * (IDC[sel] + IA[sel] * np.sin(0.25 * i * dt))[None, :])
    args, extra = cases.oracle_inputs(kw_o)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, _, _ = oracle.time_evolution(*args, len(sel), **extra)
    assert np.max(np.abs(th)) > 0.5
    assert np.max(np.abs(res.theta[:, sel, :] - th)) <= 1e-8


def test_full_size_cfg5_per_step_against_oracle():
    # BASELINE config 5 at full size: SquareArray(1000,1000) (1 998 000 junctions, 998 001 faces) with self-inductance
    # and capacitance, 64 problems: 2 048 subdomains, ~30 upper phases. Per-step parity with the oracle (whose SuperLU
    # factorisation of the 10^6-row system takes ~20 s) on four problems of the sweep
    from pyjjasim_b200 import engine
    a = pj.SquareArray(1000, 1000)
    a.set_inductance(1.0)
    a.set_capacitance(1.0)
    W, Nt, dt = 64, 8, 0.05
    base, amps = a.current_base(angle=0), np.linspace(0.5, 1.5, W)
    kw = dict(circuit=a, time_step=dt, time_step_count=Nt, external_flux=0.05, store_time_steps=[2, Nt - 1],
              store_current=False, store_voltage=False)
    res = pj.TimeEvolutionProblem(current_sources=pj.RankOneSource(base, amps), **kw).compute()
    st = engine.last_run_stats[0]
    assert st["engine"] == 3 and st["cluster_size"] == engine.subdomain_layout(a._Nf(), W, engine._sm_count(0))[2]
    sel = np.array([0, 31, 32, 63])
    th = _oracle_sub_batch(a, kw, base, amps, sel)
    assert np.max(np.abs(th)) > 0.05
    assert np.max(np.abs(res.theta[:, sel, :] - th)) <= 1e-8
    engine._tables_cache.clear()


def test_full_size_cfg2_per_step_against_oracle_with_reference_draws():
    # BASELINE config 2 at full size with the reference's own Gaussian draws injected (Nj > 500: the recycling branch of
    # time_evolution.py:533-537), all 256 temperatures, per-step parity over a short horizon
    from pyjjasim_b200 import engine
    a = pj.SquareArray(100, 100)
    W, Nt, dt = 256, 9, 0.5
    T = np.geomspace(1e-2, 1.0, W)[None, :, None]
    kw = dict(circuit=a, time_step=dt, time_step_count=Nt, external_flux=0.1, temperature=T,
              store_time_steps=[3, Nt - 1], store_current=False, store_voltage=False)
    Z = cases.replay_noise(a._Nj(), W, Nt, 11)
    res = pj.TimeEvolutionProblem(noise_replay=Z, **kw).compute()
    assert engine.last_run_stats[0]["engine"] == 3
    args, extra = cases.oracle_inputs(kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, _, _ = oracle.time_evolution(*args, W, rng=np.random.RandomState(11), **extra)
    assert np.max(np.abs(res.theta - th)) <= 1e-8


def test_running_observables_and_vortex_planes_shard_over_devices():
    # host threads, one per GPU, fill the columns of their shards (needs two devices)
    import ctypes
    from pyjjasim_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    if lib.jj_create(1, ctypes.byref(h)) != 0:
        pytest.skip("needs two CUDA devices")
    lib.jj_destroy(h)
    a = pj.SquareArray(13, 11)
    W, Nt, first, k = 24, 60, 4, 5
    kw, base, amps = _observer_case(a, W, Nt)
    one = pj.TimeEvolutionProblem(observe_interval=k, observe_first=first, store_vortex_configuration=True,
                                  store_time_steps=[10, 59], **kw).compute()
    two = pj.TimeEvolutionProblem(observe_interval=k, observe_first=first, store_vortex_configuration=True,
                                  store_time_steps=[10, 59], devices=[0, 1], **kw).compute()
    assert np.array_equal(one.get_vortex_sum(), two.get_vortex_sum())
    assert np.array_equal(one.get_vortex_configuration(), two.get_vortex_configuration())
    assert np.max(np.abs(one.get_dc_voltage() - two.get_dc_voltage())) <= 1e-9


def test_annealing_shards_over_devices():
    # two GPUs: problems never interact and the temperatures are per problem, so the sharded schedule equals the
    # single-device one exactly
    import ctypes
    from pyjjasim_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    if lib.jj_create(1, ctypes.byref(h)) != 0:
        pytest.skip("needs two CUDA devices")
    lib.jj_destroy(h)
    kw, Z, ap, (theta, n, profiles) = _anneal("anneal_small", "auto")
    kw2, Z2, ap2, (theta2, n2, profiles2) = _anneal("anneal_small", "auto", devices=[0, 1])
    assert np.array_equal(profiles, profiles2) and np.array_equal(n, n2)
    assert np.max(np.abs(theta - theta2)) <= 1e-9


def test_annealing_single_problem_and_one_step_intervals():
    # ragged sizes: one problem (padded to a quad on the device), intervals of a single step (no mobility to measure:
    # the temperature only rises), no closing-run surprises
    a = pj.SquareArray(6, 5)
    ap = pj.AnnealingProblem(a, time_step=0.5, interval_steps=1, external_flux=0.15, problem_count=1, interval_count=6,
                             vortex_mobility=0.01, start_T=0.2, T_factor=1.5, noise_seed=3)
    theta, n, prof = ap.anneal()
    assert theta.shape == (a._Nj(), 1) and n.shape == (a._Nf(), 1) and prof.shape == (6, 1)
    assert np.allclose(prof[:, 0], 0.2 * 1.5 ** np.arange(1, 7))
    assert np.array_equal(n, oracle.vortex_configuration(a.get_cycle_matrix(), theta))
    assert np.max(np.abs(a.get_cycle_matrix() @ theta + 2 * np.pi * 0.15)) < 1e-10


def test_full_length_cfg1_iv_curve_against_oracle():
    # BASELINE config 1 at full length: SquareArray(20,20), 32 bias currents, f = 0, T = 0, dt = 0.05, 10 000 steps
    # (examples/time_evolution_example_2_IV_curve.py). Non-chaotic: the phases of the running problems reach ~1e3 rad
    # and still agree with the oracle to 1e-9 relative; the DC voltages (the IV curve) agree to 1e-10
    a = pj.SquareArray(20, 20)
    W, Nt, dt = 32, 10000, 0.05
    Is = a.current_base(angle=0)[:, None, None] * np.linspace(0, 2, W)[None, :, None]
    kw = dict(circuit=a, time_step=dt, time_step_count=Nt, current_sources=Is,
              store_time_steps=[Nt // 3, Nt - 1], store_current=False, store_voltage=False)
    res = pj.TimeEvolutionProblem(**kw).compute()
    args, extra = cases.oracle_inputs(kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, _, _ = oracle.time_evolution(*args, W, **extra)
    scale = max(1.0, float(np.max(np.abs(th))))
    assert scale > 100.0                                   # the problems above the critical current do run
    assert np.max(np.abs(res.theta - th)) <= 1e-9 * scale
    span = (Nt - 1 - Nt // 3) * dt
    V_dev = (res.theta[:, :, 1] - res.theta[:, :, 0]) / span
    V_ora = (th[:, :, 1] - th[:, :, 0]) / span
    assert np.max(np.abs(V_dev - V_ora)) <= 1e-10
    # the IV curve itself: zero voltage below the array's critical current, ohmic far above it
    Vmean = (a.current_base(angle=0) @ V_dev) / np.sum(a.current_base(angle=0) ** 2)
    assert abs(Vmean[4]) < 1e-6 and Vmean[-1] > 1.5


def test_long_noisy_run_statistics_match_oracle():
    # chaotic regime (f = 0.1, T > 0): trajectories cannot be compared step by step with different generators
    # (device Philox vs the reference's MT19937), so compare what the north star names: time-averaged vortex counts
    # and DC voltages over a long run, per temperature, within the ensemble's own sampling noise
    a = pj.SquareArray(24, 24)
    W, Nt, dt = 24, 3000, 0.5
    T = np.repeat(np.array([0.05, 0.2, 0.6]), W // 3)[None, :, None]
    store = np.arange(Nt // 3, Nt, 20)
    kw = dict(circuit=a, time_step=dt, time_step_count=Nt, external_flux=0.1, temperature=T,
              store_time_steps=store, store_current=False, store_voltage=False)
    res = pj.TimeEvolutionProblem(noise_seed=77, **kw).compute()
    args, extra = cases.oracle_inputs(kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, _, _ = oracle.time_evolution(*args, W, rng=np.random.RandomState(5), **extra)
    A = a.get_cycle_matrix()

    def stats(theta):
        n = oracle.vortex_configuration(A, theta)                       # (Nf, W, K)
        dens = np.abs(n).sum(axis=0).mean(axis=1) / A.shape[0]           # vortices + antivortices per face
        net = n.sum(axis=0).mean(axis=1) / A.shape[0]                    # net vorticity per face (-> f)
        v = np.abs(theta[:, :, -1] - theta[:, :, 0]).mean(axis=0) / ((store[-1] - store[0]) * dt)
        return dens.reshape(3, -1), net.reshape(3, -1), v.reshape(3, -1)
    for got, want, name in zip(stats(res.theta), stats(th), ("density", "net vorticity", "mean |V|")):
        for t in range(3):
            g, w_ = got[t], want[t]
            sigma = np.sqrt(g.var() / g.size + w_.var() / w_.size) + 1e-3 * max(abs(w_.mean()), 1e-3) + 1e-4
            assert abs(g.mean() - w_.mean()) <= 5 * sigma, (name, t, g.mean(), w_.mean(), sigma)
    # hotter ensembles hold more vortex-antivortex pairs
    d = stats(res.theta)[0].mean(axis=1)
    assert d[2] > d[0]


def test_result_planes_come_from_the_pinned_pool_and_are_reused():
    # stored planes are fetched into page-locked blocks that go back to a pool when the result is dropped
    from pyjjasim_b200 import engine
    kw, _ = cases.build("sq_frustrated", pj)
    kw.update(store_voltage=False, store_current=False)       # (voltage trimming copies the planes with np.delete)
    r1 = pj.TimeEvolutionProblem(**kw).compute()
    ref = r1.theta.copy()
    base = r1.theta
    while getattr(base, "base", None) is not None:
        base = base.base
    assert isinstance(base, engine._PinnedBlock)
    ptr = base._ptr
    del base, r1
    import gc
    gc.collect()
    r2 = pj.TimeEvolutionProblem(**kw).compute()
    b2 = r2.theta
    while getattr(b2, "base", None) is not None:
        b2 = b2.base
    assert isinstance(b2, engine._PinnedBlock)
    assert np.array_equal(r2.theta, ref)
    assert engine._pinned._in_use > 0 and b2._ptr == ptr
