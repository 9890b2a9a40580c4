"""
Stationary states of the annealed problems: the closing step of ``AnnealingProblem.compute``
(reference: time_evolution.py:1185-1190 hands every annealed vortex configuration to ``StaticProblem.compute``,
static_problem.py:559-610, i.e. the London approximation static_problem.py:1378-1392 as initial guess followed by the
Newton iteration static_problem.py:1400-1512).

This is host code, like the reference's, and outside the time-evolution hot path: one sparse factorisation per problem
and Newton iteration (the Jacobian sandwich A (L + diag(1/q)) A^T changes with the phases of every problem). What is
shared is done once for all problems: the London approximation is one multi-right-hand-side solve, the error norms
and target checks are evaluated column-wise, and problems that have stopped drop out of the iteration.

The stopping rule and the status are the reference's (0 converged onto the target vortex configuration, 1 diverged or
converged elsewhere, 2 iteration limit), so scripts that keep the status-0 results behave the same.
"""
import numpy as np
import scipy.sparse
import scipy.sparse.csgraph
import scipy.sparse.linalg

TOL, MAXITER = 1e-10, 30           # reference: static_problem.py:25-27
TWO_PI = 2.0 * np.pi


def _spectral_norm(B):
    """||B||_2 of a sparse incidence matrix (cached by the caller); a 1-row matrix has no eigsh"""
    G = (B @ B.T).astype(np.double)
    if G.shape[0] == 1:
        return float(np.sqrt(G.toarray()[0, 0]))
    return float(np.sqrt(scipy.sparse.linalg.eigsh(G, 1, maxiter=1000, which="LA")[0][0]))


class _CircuitOps:
    """Matrices, norms and cached factorisations of one circuit."""

    def __init__(self, circuit):
        self.A = scipy.sparse.csr_matrix(circuit.get_cycle_matrix()).astype(np.double)
        self.M = scipy.sparse.csr_matrix(circuit.get_cut_matrix()).astype(np.double)
        self.L = scipy.sparse.csr_matrix(circuit._L())
        self.Ic = np.asarray(circuit._Ic(), dtype=np.double)
        self.A_norm, self.M_norm = _spectral_norm(self.A), _spectral_norm(self.M)
        self.Nf, self.Nj = self.A.shape

    def kirchhoff_error(self, I, Is, Is_norm):
        """normalised residual of M (I - Is) = 0 per problem (reference: static_problem.py:1108-1117)"""
        b = np.linalg.norm(self.M @ (I - Is), axis=0)
        scale = self.M_norm * (Is_norm + np.linalg.norm(I, axis=0))
        return np.where(np.abs(scale) < 1e-20, np.finfo(float).eps, b / np.where(scale == 0, 1.0, scale))

    def winding_error(self, theta, I, df):
        """normalised residual of A (theta + L I) + df = 0 per problem (reference: static_problem.py:1119-1130)"""
        rms = lambda x: np.linalg.norm(x, axis=0) / np.sqrt(x.shape[0])
        LI = self.L @ I
        scale = self.A_norm * (rms(theta) + rms(LI)) + rms(df)
        r = rms(df + self.A @ (theta + LI))
        return np.where(np.abs(scale) < 1e-20, np.finfo(float).eps, r / np.where(scale == 0, 1.0, scale))


def integral_cycle_solve(A, b):
    """
    An INTEGER solution Z (Nj, W) of A Z = b for integer b (Nf, W): the multiples of 2 pi that move phases from one
    phase zone to another without touching Kirchhoff's current law (the reference has a graph routine for this,
    josephson_circuit.py:782-800; any integral solution serves, they differ by integer vectors of the cut space,
    which change neither sin(theta) nor A theta).

    The faces and the outside of a planar circuit form the dual graph, whose edges are the junctions. On a spanning
    tree of it rooted at the outside, the junction between a face and its parent carries what its subtree needs:
    processed leaves first, every equation has exactly one unknown left, with coefficient +-1.
    """
    A = scipy.sparse.csc_matrix(A)
    Nf, Nj = A.shape
    b = np.asarray(b)
    Z = np.zeros((Nj, b.shape[1]), dtype=np.int64)
    if Nf == 0 or not b.any():
        return Z
    cnt = np.diff(A.indptr)
    OUT = Nf                                              # the outer face
    f1 = np.where(cnt >= 1, A.indices[np.minimum(A.indptr[:-1], A.indices.size - 1)], OUT)
    f2 = np.where(cnt >= 2, A.indices[np.minimum(A.indptr[:-1] + 1, A.indices.size - 1)], OUT)
    keep = cnt >= 1
    jj = np.flatnonzero(keep)
    dual = scipy.sparse.coo_matrix((jj + 1, (f1[keep], f2[keep])), shape=(Nf + 1, Nf + 1)).tocsr()   # value: junction + 1
    order, pred = scipy.sparse.csgraph.breadth_first_order(dual + dual.T, OUT, directed=False)
    if order.size != Nf + 1:
        raise ValueError("the faces of the circuit are not all connected to its outside")
    sym = (dual + dual.T).tocsr()         # (parallel junctions between two faces add up; any one of them serves: pick below)
    Acsr = scipy.sparse.csr_matrix(A)
    parent_j = np.full(Nf, -1, dtype=np.int64)
    need = b.astype(np.int64).copy()                      # what each face still has to receive
    for u in order[:0:-1]:                                # leaves first, the root (outside) excluded
        p = pred[u]
        # a junction shared by u and its parent
        row = Acsr.indices[Acsr.indptr[u]:Acsr.indptr[u + 1]]
        cand = row[(f1[row] == p) | (f2[row] == p)] if p != OUT else row[cnt[row] == 1]
        j = int(cand[0])
        parent_j[u] = j
        a_uj = Acsr[u, j]
        Z[j] = need[u] * int(np.sign(a_uj))               # a_uj = +-1: exact
        if p != OUT:
            need[p] -= int(Acsr[p, j]) * Z[j]
    return Z


def london_approximation(circuit, f, n, Is, ops=None):
    """
    theta0 (Nj, W) = Ic^-1 (A^T (A (Ic^-1 + L) A^T)^-1 (2 pi (n - f) - A (Ic^-1 + L) Is) + Is) for all problems at once
    (reference: static_problem.py:1378-1392), moved to phase zone 0 as StaticProblem.approximate does (:550-557).
    f, n: (Nf, W); Is: (Nj, W).
    """
    ops = ops or _CircuitOps(circuit)
    inv_Ic = np.where(np.abs(ops.Ic) > 1e-12, 1.0 / np.where(ops.Ic == 0, 1.0, ops.Ic), ops.Ic)
    if np.all(np.abs(ops.Ic) < 1e-12):
        return np.zeros((ops.Nj, np.shape(n)[1]))
    D = scipy.sparse.diags(inv_Ic) + ops.L
    solve = scipy.sparse.linalg.factorized((ops.A @ D @ ops.A.T).tocsc())
    rhs = TWO_PI * (n - f) - ops.A @ (D @ Is)
    j = np.column_stack([solve(np.ascontiguousarray(rhs[:, w])) for w in range(rhs.shape[1])])
    theta = inv_Ic[:, None] * (ops.A.T @ j + Is)
    # the approximation lives in the phase zone z = n; the Newton iteration works in zone 0 (static_problem.py:555-556)
    return theta - TWO_PI * integral_cycle_solve(ops.A, np.asarray(n))


def newton_stationary_states(circuit, theta0, Is, f, n, cpr, tol=TOL, maxiter=MAXITER, ops=None):
    """
    Newton iteration for the stationary states of W problems (reference: static_problem.py:1400-1512, with z = 0,
    stop_as_residual_increases=True, stop_if_not_target_n=False as AnnealingProblem uses it).

    theta0, Is : (Nj, W);  f, n : (Nf, W);  cpr : CurrentPhaseRelation.
    Returns theta (Nj, W), status (W,) int, info dict(iterations (W,), error (W,), on_target (W,) bool).
    """
    ops = ops or _CircuitOps(circuit)
    A, L = ops.A, ops.L
    W = theta0.shape[1]
    Ic = ops.Ic[:, None]
    theta = np.array(theta0, dtype=np.double)
    df = TWO_PI * np.asarray(f, dtype=np.double)
    LIs = L @ Is
    Is_norm = np.linalg.norm(Is, axis=0)
    target = -np.asarray(n)

    def assess(cols):
        I = cpr.eval(Ic, theta[:, cols])
        err = np.maximum(ops.kirchhoff_error(I, Is[:, cols], Is_norm[cols]), ops.winding_error(theta[:, cols], I, df[:, cols]))
        hit = np.all(A @ np.round(theta[:, cols] / TWO_PI) == target[:, cols], axis=0)
        return I, err, hit

    everyone = np.arange(W)
    I, error, on_target = assess(everyone)
    history = [error.copy()]                      # error of every problem after each iteration (frozen once it stops)
    iterations = np.zeros(W, dtype=int)
    active = np.ones(W, dtype=bool)
    for it in range(maxiter):
        # (the reference compares the error after pass `it` with the one four passes earlier, static_problem.py:1505-1506)
        prev = history[it - 4] if it >= 4 else np.full(W, np.inf)
        active &= ~((error < tol) | ((error > 0.5) & (it > 5)) | (error > prev))
        cols = np.flatnonzero(active)
        if cols.size == 0:
            break
        q = cpr.d_eval(Ic, theta[:, cols])
        q = np.where(np.abs(q) < 0.1 * tol, 0.1 * tol, q)
        y = (I[:, cols] - Is[:, cols]) / q
        rhs = A @ (theta[:, cols] - y - LIs[:, cols]) + df[:, cols]
        for k, w in enumerate(cols):
            S = L + scipy.sparse.diags(1.0 / q[:, k])
            j = scipy.sparse.linalg.spsolve((A @ S @ A.T).tocsc(), rhs[:, k])
            if not np.all(np.isfinite(j)):
                theta[:, w] += 1e10               # (the reference's way of flagging a singular step: it diverges next)
            else:
                theta[:, w] -= y[:, k] + (A.T @ j) / q[:, k]
        I_new, err_new, hit_new = assess(cols)
        I[:, cols], error[cols], on_target[cols] = I_new, err_new, hit_new
        iterations[cols] += 1
        history.append(error.copy())
    converged = error < tol
    status = (~(converged & on_target)).astype(int) + 2 * (iterations >= maxiter).astype(int)
    return theta, status, dict(iterations=iterations, error=error, on_target=on_target)


class AnnealedConfiguration:
    """
    One annealed problem after the static polish: the stationary phases found by the Newton iteration for the annealed
    vortex configuration (status 0), or the iteration's last iterate otherwise; ``annealed_theta`` keeps the phases the
    time evolution ended with. Offers the getters of the reference's StaticConfiguration that annealing scripts read.
    """

    def __init__(self, circuit, theta, n, external_flux, current_sources, cpr, annealed_theta=None, error=None):
        self.circuit, self.theta, self.n = circuit, theta, n
        self.external_flux, self.current_sources, self.current_phase_relation = external_flux, current_sources, cpr
        self.annealed_theta, self.error = annealed_theta, error

    def get_circuit(self):
        return self.circuit

    def get_theta(self):
        return self.theta

    def get_n(self):
        return self.n

    get_vortex_configuration = get_n

    def get_current(self):
        return self.current_phase_relation.eval(self.circuit._Ic(), self.theta)

    def get_phase(self):
        c = self.circuit
        return c.Msq_solve((c.get_cut_matrix() @ self.theta)[:, None])[:, 0]

    def get_cycle_current(self):
        c = self.circuit
        Is = np.broadcast_to(self.current_sources, (c._Nj(),))
        return c.Asq_solve((c.get_cycle_matrix() @ (self.get_current() - Is))[:, None])[:, 0]

    def get_flux(self):
        c = self.circuit
        return np.broadcast_to(self.external_flux, (c._Nf(),)) + c.get_cycle_matrix() @ (c._L() @ self.get_current()) / TWO_PI

    def get_josephson_energy(self):
        return self.current_phase_relation.i_eval(self.circuit._Ic(), self.theta)

    def get_magnetic_energy(self):
        I = self.get_current()
        return 0.5 * I * (self.circuit._L() @ I)

    def get_energy(self):
        return self.get_josephson_energy() + self.get_magnetic_energy()

    def get_error(self):
        return self.error
