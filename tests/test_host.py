"""CPU tests of the host logic: circuit construction, ordering, solve program, input classification, API."""
import os
import re

import numpy as np
import pytest
import scipy.sparse
import scipy.sparse.linalg

import pyjjasim_b200 as pj
from pyjjasim_b200 import sources
from pyjjasim_b200.current_phase_relation import harmonics
from pyjjasim_b200.engine import CircuitTables, shard_bounds
from pyjjasim_b200.factor import apply_program_host, build_solve_program, system_matrix
from pyjjasim_b200.ordering import nested_dissection
from tests import cases
from tests.golden.make_golden import matrix_digest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- circuits (sizes table)
@pytest.mark.parametrize("ctor,args,Nn,Nj,Nf", [
    (pj.SquareArray, (20, 20), 400, 760, 361),
    (pj.SquareArray, (100, 100), 10000, 19800, 9801),
    (pj.HoneycombArray, (20, 20), 1600, 2340, 741),
    (pj.TriangularArray, (5, 4), 40, 95, 56),
])
def test_lattice_sizes(ctor, args, Nn, Nj, Nf):
    a = ctor(*args)
    assert (a._Nn(), a._Nj(), a._Nf()) == (Nn, Nj, Nf)
    A, M = a.get_cycle_matrix(), a.get_cut_matrix()
    assert (A @ M.T).nnz == 0 or np.max(np.abs((A @ M.T).data)) == 0      # cycles are divergence free
    assert np.all(np.abs(A.data) == 1)


def test_cycle_matrices_match_reference_digests(golden_dir):
    for name in cases.CASES:
        kw, _ = cases.build(name, pj)
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        assert matrix_digest(kw["circuit"].get_cycle_matrix()) == str(g["A_digest"])


def test_square_array_conventions():
    a = pj.SquareArray(4, 3)
    n1, n2 = a.get_junction_nodes()
    assert list(n1[:3]) == [0, 1, 2] and list(n2[:3]) == [1, 2, 3]        # horizontal junctions first, row-major
    base = a.current_base(angle=0)
    assert np.allclose(base[:9], 1.0) and np.allclose(base[9:], 0.0, atol=1e-12)
    assert pj.HoneycombArray(3, 3).get_cycle_matrix().getnnz(axis=1).max() == 6
    assert np.array_equal(np.diff(pj.SQUID().get_cycle_matrix().tocsr().indptr), [4])


def test_graph_errors():
    with pytest.raises(pj.SelfLoopError):
        pj.EmbeddedGraph([0, 1], [0, 0], [0], [0])
    with pytest.raises(pj.NonSimpleError):
        pj.EmbeddedGraph([0, 1], [0, 0], [0, 1], [1, 0])
    with pytest.raises(pj.NotSingleComponentError):
        pj.EmbeddedGraph([0, 1, 2, 3], [0, 0, 1, 1], [0, 2], [1, 3], require_single_component=True)
    with pytest.raises(pj.NotPlanarEmbeddingError):
        pj.Circuit(pj.EmbeddedGraph([0, 1, 1, 0], [0, 1, 0, 1], [0, 2, 0, 1], [1, 3, 2, 3]))
    with pytest.raises(ValueError):
        pj.SquareArray(3, 3).set_resistance(0.0)
    with pytest.raises(ValueError):
        pj.SquareArray(3, 3).set_capacitance(-1.0)


# ---------------------------------------------------------------- ordering and solve program
@pytest.mark.parametrize("ctor,args,leaf", [(pj.SquareArray, (23, 17), 8), (pj.HoneycombArray, (9, 7), 6),
                                            (pj.TriangularArray, (8, 6), 16), (pj.SquareArray, (40, 40), 4)])
def test_solve_program_matches_direct_solve(ctor, args, leaf):
    a = ctor(*args)
    rng = np.random.RandomState(1)
    a.set_resistance(0.5 + rng.rand(a._Nj()))
    a.set_inductance(0.1 * rng.rand(a._Nj()))
    dt = 0.05
    S = system_matrix(a.get_cycle_matrix(), a._L(), 1 / (dt * a._R()), a._C() / dt ** 2)
    cx, cy = a.get_face_centroids()
    perm, bptr, height = nested_dissection(S, cx, cy, leaf_size=leaf)
    assert sorted(perm) == list(range(a._Nf())) and bptr[-1] == a._Nf()
    # blocks of equal height must not be coupled (they are processed concurrently)
    Sp = scipy.sparse.coo_matrix(S.tocsr()[perm][:, perm])
    blk = np.repeat(np.arange(len(bptr) - 1), np.diff(bptr))
    cross = (blk[Sp.row] != blk[Sp.col]) & (height[blk[Sp.row]] == height[blk[Sp.col]])
    assert not np.any(cross)
    prog = build_solve_program(S, cx, cy, leaf_size=leaf)
    b = rng.randn(a._Nf(), 3)
    J = apply_program_host(prog, b)
    Jref = scipy.sparse.linalg.spsolve(S.tocsc(), b)
    assert np.max(np.abs(J - Jref)) <= 1e-12 * np.max(np.abs(Jref))
    for sw in prog.sweeps.values():
        assert np.all(sw["tile_nrows"] * sw["tile_lpr"] <= 32)
        assert sw["group_ptr"][-1] == len(sw["tile_row0"])
    assert prog.stats["levels_fwd"] == 2 * prog.stats["height"] + 1


@pytest.mark.parametrize("ctor,args,d,NG", [(pj.SquareArray, (23, 17), 2, 4), (pj.HoneycombArray, (9, 7), 1, 2),
                                            (pj.SquareArray, (40, 37), 3, 4), (pj.SquareArray, (12, 10), 0, 1),
                                            (pj.TriangularArray, (8, 6), 2, 8)])
def test_subdomain_plan_matches_direct_solve(ctor, args, d, NG):
    # the subdomain engine's plan (local sweeps + dense top inverse), interpreted on the host
    from pyjjasim_b200.subdomain import apply_subdomain_plan_host
    a = ctor(*args)
    rng = np.random.RandomState(1)
    a.set_resistance(0.5 + rng.rand(a._Nj()))
    a.set_inductance(0.1 * rng.rand(a._Nj()))
    tab = CircuitTables(a, 0.05)
    plan = tab.subdomain_plan(d, NG)
    assert plan.P == 1 << d and plan.n_top + int(plan.n_loc.sum()) == a._Nf()
    # every junction is owned once, every face entry appears in exactly one list
    assert plan.junc_ptr[-1] == a._Nj() and np.sum(plan.face_ell_j >= 0) == a.get_cycle_matrix().nnz
    S = system_matrix(a.get_cycle_matrix(), a._L(), tab.Rv, tab.Cv)
    b = rng.randn(a._Nf(), plan.PC)
    Jp = apply_subdomain_plan_host(plan, b[tab.perm])
    J = np.empty_like(Jp)
    J[tab.perm] = Jp
    Jref = scipy.sparse.linalg.spsolve(S.tocsc(), b)
    assert np.max(np.abs(J - Jref)) <= 1e-12 * np.max(np.abs(Jref))


@pytest.mark.parametrize("ctor,args,n_parts,NG", [(pj.SquareArray, (23, 17), 3, 2), (pj.HoneycombArray, (9, 7), 5, 1),
                                                  (pj.SquareArray, (40, 37), 6, 4), (pj.SquareArray, (20, 20), 1, 4)])
def test_subdomain_plan_with_uneven_dissection(ctor, args, n_parts, NG):
    # orderings made with n_parts subtrees (one per SM and problem chunk, rarely a power of two)
    from pyjjasim_b200.subdomain import apply_subdomain_plan_host
    a = ctor(*args)
    rng = np.random.RandomState(4)
    a.set_resistance(0.5 + rng.rand(a._Nj()))
    tab = CircuitTables(a, 0.05, n_parts=n_parts)
    part = np.asarray(tab.factor.blk_part)
    sizes = np.diff(tab.factor.bptr)
    rows = np.array([sizes[part == k].sum() for k in range(part.max() + 1)])
    # (cuts run along lattice lines: on a tiny lattice a domain can get too small to be cut into all its parts)
    n_got = part.max() + 1
    assert (n_got == n_parts or (a._Nf() < 200 and n_parts - 1 <= n_got <= n_parts)) and (a._Nf() < 500 or rows.min() >= 0.6 * rows.max())
    plan = tab.subdomain_plan(None, NG)
    assert plan.P == n_got
    S = system_matrix(a.get_cycle_matrix(), a._L(), tab.Rv, tab.Cv)
    b = rng.randn(a._Nf(), plan.PC)
    Jp = apply_subdomain_plan_host(plan, b[tab.perm])
    J = np.empty_like(Jp)
    J[tab.perm] = Jp
    Jref = scipy.sparse.linalg.spsolve(S.tocsc(), b)
    assert np.max(np.abs(J - Jref)) <= 1e-12 * np.max(np.abs(Jref))
    # the streaming program of the same ordering is still a valid solve
    J2 = apply_program_host(tab.program, b)
    assert np.max(np.abs(J2 - Jref)) <= 1e-12 * np.max(np.abs(Jref))


def test_circuit_tables_permutation_consistency():
    a = pj.HoneycombArray(6, 5)
    tab = CircuitTables(a, 0.05)
    A = a.get_cycle_matrix().tocsr()
    # CSR rows in permuted order reproduce A[perm]
    rebuilt = scipy.sparse.csr_matrix((tab.face_sign.astype(float), tab.face_junc, tab.face_ptr), shape=A.shape)
    assert abs(rebuilt - A[tab.perm]).nnz == 0
    # junction -> faces table is the transpose
    for j in range(a._Nj()):
        for k in range(2):
            f = tab.junc_face[j, k]
            if f >= 0:
                assert A[tab.perm[f], j] == tab.junc_sign[j, k]
    assert np.sum(tab.junc_face >= 0) == A.nnz
    assert np.sum(np.all(tab.junc_face < 0, axis=1)) == 2        # the honeycomb's two face-less corner junctions


# ---------------------------------------------------------------- inputs
def test_source_classification():
    N, W, Nt = 12, 5, 9
    rng = np.random.RandomState(0)
    assert sources.classify_source(0.0, N, W, Nt).kind == sources.ZERO
    assert sources.classify_source(1e-9, N, W, Nt).kind == sources.ZERO           # reference quirk Q5 (allclose)
    s = sources.classify_source(0.3, N, W, Nt)
    assert s.kind == sources.RANK1 and s.static and np.array_equal(s.value(4), np.full((N, W), 0.3))
    amp = np.linspace(0, 2, W)
    base = rng.randn(N)
    x = base[:, None, None] * amp[None, :, None]
    s = sources.classify_source(x, N, W, Nt)
    assert s.kind == sources.RANK1 and s.static and np.allclose(s.value(0), x[:, :, 0], rtol=1e-15)
    xt = rng.randn(1, W, Nt)
    s = sources.classify_source(xt, N, W, Nt)
    assert s.kind == sources.RANK1 and not s.static and np.array_equal(s.amp_chunk(2, 5), xt[0, :, 2:5].T)
    xd = rng.randn(N, W, 1)
    s = sources.classify_source(xd, N, W, Nt)
    assert s.kind == sources.DENSE and s.static and np.array_equal(s.dense_chunk(0, 1)[0], xd[:, :, 0])
    fn = lambda i: base[:, None] * (amp + np.sin(0.1 * i))
    s = sources.classify_source(fn, N, W, Nt)
    assert s.kind == sources.RANK1 and not s.static
    assert np.allclose(s.base[:, None] * s.amp_chunk(3, 4)[0][None, :], fn(3), rtol=1e-15)
    s = sources.classify_source(lambda i: rng.randn(N, W), N, W, Nt)
    assert s.kind == sources.DENSE
    r1 = pj.RankOneSource(base, lambda i: amp * i)
    assert r1.problem_count == W and np.array_equal(r1(2), base[:, None] * (2 * amp)[None, :])
    with pytest.raises(ValueError):
        sources.classify_source(np.zeros((N, W)), N, W, Nt)                        # quirk Q6: 2-D input is misread


def _refstyle_cases(pkg, mk_problem):
    """(problem, expected kinds) for reference-style problem objects: constants must not look time dependent just because
    the reference holds them as (N, W, Nt) broadcast views (time_evolution.py:117-131)"""
    a = pkg.SquareArray(9, 8)
    W, Nt = 6, 400
    base = a.current_base(angle=0)
    amps = np.linspace(0.2, 1.4, W)
    T = np.linspace(0.0, 0.3, W)[None, :, None]
    p1 = mk_problem(a, time_step=0.05, time_step_count=Nt, external_flux=0.1, temperature=T,
                    current_sources=base[:, None, None] * amps[None, :, None])
    p2 = mk_problem(a, time_step=0.05, time_step_count=Nt, current_sources=lambda i: base[:, None] * (amps + 0.1 * np.sin(0.3 * i)),
                    voltage_sources=0.0, external_flux=np.linspace(0, 0.2, Nt)[None, None, :] * np.ones((1, W, 1)))
    return a, (p1, dict(Is=(sources.RANK1, True), f=(sources.RANK1, True), Vs=(sources.ZERO, True), T=(sources.RANK1, True))), \
              (p2, dict(Is=(sources.RANK1, False), f=(sources.RANK1, False), Vs=(sources.ZERO, True), T=(sources.ZERO, True)))


def _check_refstyle_classification(pkg, mk_problem, wrap):
    import tracemalloc
    from pyjjasim_b200 import engine
    a, *probs = _refstyle_cases(pkg, mk_problem)
    for prob, expect in probs:
        assert not hasattr(prob, "_raw_sources")
        tab = engine.CircuitTables(wrap(a), 0.05)
        tracemalloc.start()
        specs = engine._classify_all(prob, tab)
        peak = tracemalloc.get_traced_memory()[1]
        tracemalloc.stop()
        for name, (kind, static) in expect.items():
            assert specs[name].kind == kind and bool(specs[name].static) == static, (name, specs[name].kind, specs[name].static)
        # nothing of size N x W x Nt is materialised to classify a broadcast view (advice r1: ~1 TB on cfg2 at Nt = 1e5)
        assert peak < 0.25 * a._Nj() * prob.get_problem_count() * prob._Nt() * 8, peak


def test_reference_style_problem_objects_are_classified_by_their_frozen_flags():
    from tests.refstyle import RefStyleProblem, RefStyleCircuit
    mk = lambda a, **kw: RefStyleProblem(RefStyleCircuit(a), current_phase_relation=pj.DefaultCPR(), **kw)
    _check_refstyle_classification(pj, mk, RefStyleCircuit)


def test_the_reference_own_problem_class_binds_to_the_device_core_classification():
    # the real thing, where the reference tree exists (build container): the unmodified reference's
    # TimeEvolutionProblem / SquareArray objects, as INTEGRATION.md binds them
    from oracle.ref_harness import reference_available, import_reference
    if not reference_available():
        pytest.skip("reference tree not present")
    ref = import_reference()
    from tests.refstyle import RefStyleProblem, RefStyleCircuit
    _check_refstyle_classification(ref, lambda a, **kw: ref.TimeEvolutionProblem(a, **kw), lambda a: a)
    # and the stand-in used on the GPU box holds its inputs in the same form as the real class
    a = ref.SquareArray(5, 4)
    kw = dict(time_step=0.1, time_step_count=7, external_flux=0.2, temperature=np.ones((1, 3, 1)))
    real, fake = ref.TimeEvolutionProblem(a, **kw), RefStyleProblem(RefStyleCircuit(a), **kw)
    for attr in ("external_flux", "current_sources", "voltage_sources", "temperature", "config_at_minus_1", "config_at_minus_2"):
        x, y = getattr(real, attr), getattr(fake, attr)
        assert x.shape == y.shape and x.strides == y.strides and np.array_equal(x, y), attr
    for flag in ("_f_is_timedep", "_Is_is_timedep", "_Vs_is_timedep", "_T_is_timedep"):
        assert getattr(real, flag) == getattr(fake, flag)
    assert real.get_problem_count() == fake.get_problem_count() and real._Nt() == fake._Nt() and real._dt() == fake._dt()


def test_closed_form_amplitudes_are_recovered_from_plain_callables(monkeypatch):
    # a reference-style callable Is(i) -> (N, W) of the form base x (constant + ramp + DC/AC drive) is evaluated at a few
    # dozen steps instead of at every step; anything else keeps the per-step evaluation; a model that stops holding is
    # retired at the next check and the chunk is evaluated exactly
    N, W, Nt = 30, 17, 60000
    rng = np.random.RandomState(1)
    base, IDC = rng.randn(N), np.linspace(0, 2, W)
    j0 = int(np.argmax(np.abs(base)))
    calls = [0]

    def counted(f):
        def g(i):
            calls[0] += 1
            return f(i)
        return g
    for amp in (lambda i: IDC + np.sin(0.0125 * i), lambda i: 0.3 * IDC + (1 + IDC) * np.cos(2.9 * i + 0.3) + 1e-5 * i,
                lambda i: 0.5 + 1e-4 * i * np.ones(W), lambda i: IDC + 0.0 * i):
        calls[0] = 0
        s = sources.classify_source(counted(lambda i, a=amp: base[:, None] * a(i)[None, :]), N, W, Nt)
        assert s.kind == sources.RANK1 and s.model is not None
        for i0 in (0, 31000, Nt - 500):
            tab = s.amp_chunk(i0, i0 + 500) * s.base[j0]
            want = np.stack([amp(i) for i in range(i0, i0 + 500)]) * base[j0]
            assert np.max(np.abs(tab - want)) <= 1e-9 * np.max(np.abs(want))
        assert s.model is not None and calls[0] < 120            # not 1500 evaluations
    # a chirp is not one of the closed forms: evaluated per step
    s = sources.classify_source(lambda i: base[:, None] * (1 + np.sin(1e-6 * i * i)) * np.ones((1, W)), N, W, Nt)
    assert s.kind == sources.RANK1 and s.model is None
    # a change that the far samples of the fit see: no model
    late = lambda i: base[:, None] * (IDC + np.sin(0.01 * i) * (1.0 if i < 30000 else 1.5))[None, :]
    assert sources.classify_source(late, N, W, Nt).model is None
    # the closed form stops holding between steps 20 000 and 26 000 (none of the fit's samples falls there): the model is
    # retired at the next check and the chunk is evaluated exactly
    late = lambda i: base[:, None] * (IDC + np.sin(0.01 * i) * (1.5 if 20000 <= i < 26000 else 1.0))[None, :]
    s = sources.classify_source(late, N, W, Nt)
    assert s.model is not None
    s.model.check_every = 256
    assert np.allclose(s.amp_chunk(19000, 19900)[5] * s.base[j0], (IDC + np.sin(0.01 * 19005)) * base[j0], rtol=0, atol=1e-9)
    tab = s.amp_chunk(19900, 20600)
    assert s.model is None
    assert np.allclose(tab[400] * s.base[j0], (IDC + 1.5 * np.sin(0.01 * 20300)) * base[j0], rtol=1e-14)
    monkeypatch.setenv("JJ_SOURCE_MODEL", "0")
    assert sources.classify_source(lambda i: base[:, None] * (IDC + np.sin(0.0125 * i))[None, :], N, W, Nt).model is None


@pytest.mark.parametrize("name", ["square", "honeycomb"])
def test_static_polish_matches_the_reference_static_solver(name, golden_dir):
    # London approximation in phase zone 0 and the Newton iteration against StaticProblem.approximate() / .compute() of
    # the unmodified reference (tests/golden/make_golden_static.py): same status, same iteration count, same phases up
    # to integer multiples of 2 pi (any integral solution of A Z = n moves between phase zones)
    from tests import cases
    from pyjjasim_b200.static_polish import london_approximation, newton_stationary_states, integral_cycle_solve
    ctor, args, L, f, n, Is = cases.static_cases(pj)[name]
    g = np.load(os.path.join(golden_dir, f"static_{name}.npz"))
    a = getattr(pj, ctor)(*args)
    if L:
        a.set_inductance(L)
    A = a.get_cycle_matrix()
    assert np.array_equal(A @ integral_cycle_solve(A, n), n)
    pv = lambda x: x - 2 * np.pi * np.round(x / (2 * np.pi))
    th0 = london_approximation(a, f, n, Is)
    assert np.max(np.abs(pv(th0 - g["theta0"]))) < 1e-13
    th, status, info = newton_stationary_states(a, th0, Is, f, n, pj.DefaultCPR())
    assert np.array_equal(status, g["status"]) and np.array_equal(info["iterations"], g["iterations"])
    ok = status == 0
    assert ok.sum() >= 3 and (~ok).sum() >= 1
    assert np.max(np.abs(pv(th[:, ok] - g["theta"][:, ok]))) < 1e-12
    # a converged state is a stationary state with the requested vortices
    I = np.sin(th[:, ok])
    assert np.max(np.abs(a.get_cut_matrix() @ (I - Is[:, ok]))) < 1e-9
    assert np.array_equal(-(A @ np.round(th[:, ok] / (2 * np.pi))), n[:, ok])


def test_cpr_harmonics():
    a, b = harmonics(pj.DefaultCPR())
    assert list(a) == [0, 0] and list(b) == [0, 1]
    cube = pj.CurrentPhaseRelation(lambda Ic, th: Ic * np.sin(th) ** 3, None, None)
    a, b = harmonics(cube)
    assert np.allclose(b[[1, 3]], [0.75, -0.25]) and np.allclose(a, 0) and len(b) == 4
    with pytest.raises(ValueError):
        harmonics(pj.CurrentPhaseRelation(lambda Ic, th: Ic * th, None, None))      # not periodic
    with pytest.raises(ValueError):
        harmonics(pj.CurrentPhaseRelation(lambda Ic, th: Ic ** 2 * np.sin(th), None, None))


# ---------------------------------------------------------------- API mirror (reference: time_evolution.py:94-149, 384-408)
def test_problem_constructor_semantics():
    a = pj.SquareArray(4, 4)
    Nj, Nf = a._Nj(), a._Nf()
    p = pj.TimeEvolutionProblem(a, time_step_count=10, current_sources=np.ones((Nj, 3, 1)), temperature=np.ones((1, 3, 10)))
    assert p.get_problem_count() == 3 and p._T_is_timedep and not p._Is_is_timedep
    assert p.current_sources.shape == (Nj, 3, 10) and p.external_flux.shape == (Nf, 3, 10)
    assert p.config_at_minus_1.shape == (Nj, 3) and np.all(p.config_at_minus_2 == 0)
    assert pj.TimeEvolutionProblem(a, time_step_count=10, current_sources=lambda i: np.ones((Nj, 7))).get_problem_count() == 7
    p = pj.TimeEvolutionProblem(a, time_step_count=10, store_time_steps=[2, 5])
    assert p._Nt_s() == 2 and list(np.flatnonzero(p.store_time_steps)) == [2, 5]
    mask = np.zeros(10, dtype=bool); mask[7] = True
    assert pj.TimeEvolutionProblem(a, time_step_count=10, store_time_steps=mask)._Nt_s() == 1
    with pytest.raises(ValueError, match="No output is stored"):
        pj.TimeEvolutionProblem(a, store_theta=False, store_voltage=False, store_current=False)
    with pytest.raises(ValueError, match="No output is stored"):
        pj.TimeEvolutionProblem(a, time_step_count=10, store_time_steps=np.zeros(10, dtype=bool))
    with pytest.raises(ValueError):
        pj.TimeEvolutionProblem(a, time_step_count=10, store_time_steps=[11])
    with pytest.raises(ValueError):
        pj.TimeEvolutionProblem(a, stencil_width=6)
    with pytest.raises(NotImplementedError):
        pj.TimeEvolutionProblem(a, stencil_width=4)
    assert np.array_equal(pj.TimeEvolutionProblem(a, time_step=0.1, time_step_count=3).get_time(), [0, 0.1, 0.2])


def test_result_container_semantics():
    a = pj.SquareArray(4, 4)
    Nj = a._Nj()
    p = pj.TimeEvolutionProblem(a, time_step_count=6, store_time_steps=[1, 4], store_voltage=False)
    th = np.random.RandomState(0).randn(Nj, 1, 2)
    r = pj.TimeEvolutionResult(p, th, th.copy(), None)
    assert r.voltage is None and r.get_theta().shape == (Nj, 1, 2)
    assert np.array_equal(r.get_theta([4])[:, :, 0], th[:, :, 1])
    with pytest.raises(pj.VoltageNotStored):
        r.get_voltage()
    with pytest.raises(pj.DataAtTimepointNotStored):
        r.get_theta([3])
    with pytest.raises(ValueError):
        pj.TimeEvolutionResult(p, th[:, :, :1], th, None)
    n = r.get_vortex_configuration()
    assert n.shape == (a._Nf(), 1, 2) and n.dtype.kind == "i"
    assert r.get_phase().shape == (a._Nn(), 1, 2)


def test_vortex_configuration_from_device_planes_selects_and_widens_like_the_host_path():
    # store_vortex_configuration: the result holds (K, Nf, W) int32 planes from the device (a page-locked block); the
    # getter returns the reference's (Nf, W, K) int layout for all stored steps, a subset, and a reordered subset
    a = pj.SquareArray(4, 5)
    Nj, Nf, W = a._Nj(), a._Nf(), 3
    stored = [1, 4, 5, 8]
    p = pj.TimeEvolutionProblem(a, time_step_count=10, store_time_steps=stored, store_voltage=False, store_current=False,
                                current_sources=np.zeros((Nj, W, 1)))
    rng = np.random.RandomState(4)
    th = 9.0 * rng.randn(Nj, W, len(stored))
    A = a.get_cycle_matrix()
    want = np.stack([-(A @ np.round(th[:, :, k] / (2 * np.pi))) for k in range(len(stored))], axis=2).astype(int)
    planes = np.ascontiguousarray(np.moveaxis(want, 2, 0)).astype(np.int32)
    r = pj.TimeEvolutionResult(p, th, None, None, observed=dict(n_planes=planes))
    n = r.get_vortex_configuration()
    assert n.shape == (Nf, W, len(stored)) and n.dtype == np.dtype(int) and np.array_equal(n, want)
    assert np.array_equal(r.get_vortex_configuration([4, 8]), want[:, :, [1, 3]])
    assert np.array_equal(np.reshape(r.get_vortex_configuration(5), (Nf, W)), want[:, :, 2])
    # and it agrees with the host derivation from the phases
    r_host = pj.TimeEvolutionResult(p, th, None, None)
    assert np.array_equal(r_host.get_vortex_configuration(), want)
    with pytest.raises(pj.DataAtTimepointNotStored):
        r.get_vortex_configuration([2])


def test_derived_quantities_are_batched_over_time_points_and_match_the_per_step_formulas():
    # every getter evaluates all selected time points in one batched operation; compare with the formulas of the
    # reference written out per time point (time_evolution.py:692-983)
    a = pj.SquareArray(5, 4)
    a.set_inductance(0.3)
    a.set_capacitance(0.7)
    Nj, Nf, Nn, W, Nt = a._Nj(), a._Nf(), a._Nn(), 3, 8
    rng = np.random.RandomState(3)
    Is = rng.randn(Nj, W, Nt)
    f = rng.rand(Nf, W, 1)
    p = pj.TimeEvolutionProblem(a, time_step_count=Nt, store_time_steps=[1, 4, 6], current_sources=Is, external_flux=f)
    th, I, V = 7 * rng.randn(Nj, W, 3), rng.randn(Nj, W, 3), rng.randn(Nj, W, 3)
    r = pj.TimeEvolutionResult(p, th, I, V)
    A, M, L, C, Ic = a.get_cycle_matrix(), a.get_cut_matrix(), a._L(), a._C(), a._Ic()
    for sel, planes in ((None, [0, 1, 2]), ([6, 1], [0, 2]), ([4], [1])):
        steps = [1, 4, 6] if sel is None else sorted(sel)
        want = {
            "phase": [a.Msq_solve(M @ th[:, :, k]) for k in planes],
            "vortex_configuration": [-A @ np.round(th[:, :, k] / (2 * np.pi)) for k in planes],
            "josephson_energy": [Ic[:, None] * (1 - np.cos(th[:, :, k])) for k in planes],
            "supercurrent": [Ic[:, None] * np.sin(th[:, :, k]) for k in planes],
            "cycle_current": [a.Asq_solve(A @ (I[:, :, k] - Is[:, :, t])) for k, t in zip(planes, steps)],
            "flux": [f[:, :, 0] + A @ (L @ I[:, :, k]) / (2 * np.pi) for k in planes],
            "magnetic_energy": [0.5 * L @ I[:, :, k] ** 2 for k in planes],
            "potential": [a.Msq_solve(M @ V[:, :, k]) for k in planes],
            "capacitive_energy": [0.5 * C[:, None] * V[:, :, k] ** 2 for k in planes],
        }
        for name, per_step in want.items():
            got = getattr(r, "get_" + name)(sel)
            assert got.shape == (per_step[0].shape[0], W, len(planes)), name
            assert np.allclose(got, np.stack(per_step, axis=2), rtol=1e-13, atol=1e-13), name
        assert np.allclose(r.get_energy(sel), r.get_josephson_energy(sel) + r.get_magnetic_energy(sel) + r.get_capacitive_energy(sel))
    b = pj.SquareArray(3, 3)
    r0 = pj.TimeEvolutionResult(pj.TimeEvolutionProblem(b, time_step_count=2), *(np.ones((b._Nj(), 1, 2)),) * 3)
    assert not r0.get_magnetic_energy().any() and not r0.get_capacitive_energy().any()      # no L, no C: zero
    with pytest.raises(ValueError, match="no running observables"):
        r0.get_vortex_sum()


def test_store_plan_keeps_the_helper_planes_of_the_voltage_difference():
    a = pj.SquareArray(3, 3)
    p = pj.TimeEvolutionProblem(a, time_step_count=9, store_time_steps=[0, 3, 4, 8])
    plan = pj.StorePlan(p)
    assert list(np.flatnonzero(plan.theta_mask)) == [0, 2, 3, 4, 7, 8]
    assert list(np.flatnonzero(plan.current_mask)) == [0, 3, 4, 8]        # no inductance: no helper planes for currents
    at = pj.StorePlan.positions(plan.theta_mask, plan.requested)
    assert list(at) == [2, 4, 5, 7] and list(at - 1) == [1, 3, 4, 6]       # step 0 differences against theta(-1)
    a.set_inductance(0.2)
    assert list(np.flatnonzero(pj.StorePlan(p).current_mask)) == [0, 2, 3, 4, 7, 8]
    q = pj.TimeEvolutionProblem(a, time_step_count=9, store_time_steps=[5], store_voltage=False, store_theta=False)
    plan = pj.StorePlan(q)
    assert not plan.theta_mask.any() and list(np.flatnonzero(plan.current_mask)) == [5] and not plan.fetch_theta
    o = pj.TimeEvolutionProblem(a, time_step_count=9, store_theta=False, store_voltage=False, store_current=False, observe_interval=2)
    assert not pj.StorePlan(o).theta_mask.any() and o.starts_at_rest_with_zero_phases()
    assert o.config_at_minus_2.shape == (a._Nj(), 1) and not o.config_at_minus_2.any()


def test_shard_bounds():
    assert shard_bounds(256, 8) == [0, 32, 64, 96, 128, 160, 192, 224, 256]
    assert shard_bounds(10, 4) == [0, 4, 8, 10, 10]
    b = shard_bounds(4096 + 3, 8)
    assert b[0] == 0 and b[-1] == 4099 and all(v % 4 == 0 for v in b[:-1]) and sorted(b) == b


# ---------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    import ctypes
    from pyjjasim_b200 import _lib
    header = open(os.path.join(ROOT, "include", "jjstep.h")).read()
    declared = set(re.findall(r"\b(jj_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    # struct layouts of the ctypes mirror equal the C compiler's
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "jjstep.h"\nint main(){printf("%zu %zu %zu %zu %zu", sizeof(JJSweep), sizeof(JJCircuit), sizeof(JJStats), sizeof(JJSubProgram), sizeof(JJSubdomainPlan));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [ctypes.sizeof(_lib.JJSweep), ctypes.sizeof(_lib.JJCircuit), ctypes.sizeof(_lib.JJStats),
                     ctypes.sizeof(_lib.JJSubProgram), ctypes.sizeof(_lib.JJSubdomainPlan)]


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    kw, _ = cases.build("single_problem", pj)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        pj.TimeEvolutionProblem(**kw).compute()


# ---------------------------------------------------------------- annealing caller, host side
def test_annealing_problem_mirrors_reference_host_logic():
    from tests import cases
    from oracle import oracle
    kw, _ = cases.ANNEAL_CASES["anneal_small"](pj)
    ap = pj.AnnealingProblem(**kw)
    assert ap.T.shape == (1, kw["problem_count"], 1) and np.all(ap.T == kw["start_T"])
    rng = np.random.RandomState(0)
    Nf = kw["circuit"].face_count()
    n = rng.randint(-1, 2, size=(Nf, kw["problem_count"], kw["interval_steps"]))
    mob = ap.get_vortex_mobility(n)
    assert np.array_equal(mob, oracle.vortex_mobility(n, Nf, kw["time_step"], kw["interval_count"]))
    T0 = ap.T.copy()
    ap._temperature_adjustment(mob, 3)
    upper = kw["vortex_mobility"] * ((kw["interval_count"] - 3) / kw["interval_count"]) ** 1.5
    want = np.where(mob > upper, 1 / kw["T_factor"], kw["T_factor"])
    assert np.array_equal(ap.T[0, :, 0], T0[0, :, 0] * want)
    prob = ap._problem()
    assert prob.get_problem_count() == kw["problem_count"] and prob._Nt() == kw["interval_steps"]
    assert not prob.store_current and not prob.store_voltage and not prob._T_is_timedep


def test_inputs_replaced_after_construction_are_honoured():
    # the reference's annealing loop assigns prob.temperature between compute() calls (time_evolution.py:1166); the
    # time-dependence flag stays as constructed, so the new array is read at step 0 only (:509-519)
    from pyjjasim_b200 import engine
    from pyjjasim_b200.sources import RANK1, ZERO
    a = pj.SquareArray(4, 4)
    W = 3
    prob = pj.TimeEvolutionProblem(a, time_step_count=5, temperature=0.5 * np.ones((1, W, 1)),
                                   store_current=False, store_voltage=False)
    tab = type("T", (), dict(Nj=a._Nj(), Nf=a._Nf()))()
    s = engine._classify_all(prob, tab)
    assert s["T"].kind == RANK1 and np.allclose(s["T"].amp_chunk(0, 1), 0.5)
    prob.temperature = np.array([0.1, 0.2, 0.3])[None, :, None] * np.ones((1, 1, 5))
    s = engine._classify_all(prob, tab)
    assert s["T"].kind == RANK1 and s["T"].static and np.allclose(s["T"].amp_chunk(0, 1), [[0.1, 0.2, 0.3]])
    prob.temperature = np.zeros((1, 1, 5))
    assert engine._classify_all(prob, tab)["T"].kind == ZERO


def test_subdomain_layout_rules():
    # (problem groups per chunk, chunks, subdomains) of the subdomain engine on a 148-SM device
    from pyjjasim_b200.engine import subdomain_layout
    assert subdomain_layout(9801, 256) == (4, 8, 18)          # cfg2: one (subdomain, chunk) item per block, 144 blocks
    assert subdomain_layout(361, 32) == (4, 1, 8)             # cfg1: no subdomain smaller than ~45 faces
    # larger circuits: several items per block, a power of two of subdomains (median cuts: congruent shapes share programs)
    assert subdomain_layout(65025, 512) == (4, 16, 128)       # cfg4 per GPU
    assert subdomain_layout(79401, 512) == (4, 16, 256)       # cfg3
    assert subdomain_layout(998001, 64) == (4, 2, 2048)       # cfg5: 28 items per block; 10^5 separator rows above them
    assert subdomain_layout(100, 8) == (1, 1, 2)


def test_large_subdomain_plan_falls_back_instead_of_raising():
    # subdomains too large for shared memory must not make the tables constructor fail (the streaming engine runs them)
    from pyjjasim_b200 import engine
    a = pj.SquareArray(70, 70)
    tab = engine.CircuitTables(a, 0.05, n_parts=2)            # 2 400 rows per subdomain
    assert tab.choose_subdomain(64) is None


def test_annealing_problem_input_errors_mirror_reference():
    # the reference hands current_sources straight to TimeEvolutionProblem, where a (Nj,) array is broadcast against
    # (Nj, W, Nt) along the TIME axis and fails; scalars and (Nj, 1, 1) arrays work
    a = pj.SquareArray(5, 5)
    with pytest.raises(ValueError):
        pj.AnnealingProblem(a, current_sources=np.ones(a._Nj()), problem_count=3, interval_steps=7)._problem()
    prob = pj.AnnealingProblem(a, current_sources=0.1 * np.ones((a._Nj(), 1, 1)), problem_count=3, interval_steps=7)._problem()
    assert prob.get_problem_count() == 3 and prob._Is(0).shape == (a._Nj(), 3)
    # per-face flux array
    prob = pj.AnnealingProblem(a, external_flux=np.linspace(0, 0.3, a._Nf()), problem_count=2)._problem()
    assert prob._f(0).shape == (a._Nf(), 2)
