"""
Host-side classification of the four per-step inputs of a time evolution
(current sources Is, external flux f, voltage sources Vs, temperature T).

The reference slices each input to a dense (N, W) float64 array every step
(reference: time_evolution.py:342-356, 524-531). Streaming such slices to the GPU every step would
cap throughput far below HBM speed, so the device path keeps each input in one of three forms and
evaluates it on the device:

  ZERO    value(e, w, i) = 0
  RANK1   value(e, w, i) = base[e] * amp[i][w]      amp uploaded per chunk of steps, (K, W) float64;
                                                    a time-independent input has a single row
  DENSE   value(e, w, i) = table[i][e][w]           (K, N, W) per chunk; slow path, still on device

``RankOneSource`` is a convenience callable that is also understood by the reference
(it returns the dense (N, W) slice when called), so scripts stay portable.
"""
import numpy as np

__all__ = ["RankOneSource", "SourceSpec", "classify_source", "AmplitudeModel"]

ZERO, RANK1, DENSE = 0, 1, 2
# an input counts as base[e] * amp[w] when that product reproduces it to a few units in the last place
# (the user's own array was rounded the same number of times)
_RANK1_RTOL = 2e-15


class RankOneSource:
    """
    Callable ``i -> base[:, None] * amp(i)[None, :]`` of shape (N, W).

    base : (N,) array;  amp : callable i -> (W,) array, or (W,) array (time independent).
    Passing it as current_sources / external_flux / voltage_sources / temperature lets the device path
    skip materialising (N, W) arrays on the host.
    """

    def __init__(self, base, amp, problem_count=None):
        self.base = np.ascontiguousarray(base, dtype=np.double).ravel()
        self._amp = amp
        if callable(amp):
            self.problem_count = int(np.asarray(amp(0)).size) if problem_count is None else problem_count
        else:
            self._amp_arr = np.ascontiguousarray(amp, dtype=np.double).ravel()
            self.problem_count = self._amp_arr.size

    def amp(self, i):
        if callable(self._amp):
            return np.broadcast_to(np.asarray(self._amp(i), dtype=np.double).ravel(), (self.problem_count,))
        return self._amp_arr

    def is_static(self):
        return not callable(self._amp)

    def __call__(self, i):
        return self.base[:, None] * self.amp(i)[None, :]


class AmplitudeModel:
    """
    amp(i)[w] = c[w] + l[w] i + a[w] cos(omega i) + b[w] sin(omega i): the closed forms the reference's example scripts
    drive their circuits with (a constant, a ramp, DC + AC: examples/time_evolution_example_5_giant_shapiro_steps.py:33),
    recovered from a plain callable ``Is(i) -> (N, W)`` so that the callable - which materialises an (N, W) array per
    call, 535 MB for SquareArray(256,256) x 512 problems - is evaluated at a few dozen steps instead of at every step.

    ``fit(sample)`` takes ``sample(i) -> (W,)`` (the amplitude of the rank-one input at step i), samples the first
    steps and a few far ones, and returns a model that reproduces all of them to ``RTOL`` of the amplitude scale, or
    None. While the run goes on the model is re-verified against the callable every ``check_every`` steps - an interval
    chosen from the measured cost of one call so that verification costs ~50 us per time step, between 16 and 4096
    steps - and from the first chunk of steps in which a check fails the callable is evaluated at every step again.
    What this cannot see is a transient that begins and ends between two checks (a pulse shorter than the interval on
    top of an otherwise closed-form drive): pass such inputs as arrays or RankOneSource, or set JJ_SOURCE_MODEL=0,
    which keeps the per-step evaluation. The model reproduces the callable to ~1e-10 of the amplitude scale.
    """
    RTOL = 1e-11
    NEAR = 10                # consecutive steps sampled at the start

    def __init__(self, c, l, a, b, omega, check_every=2048):
        self.c, self.l, self.a, self.b, self.omega = c, l, a, b, float(omega)
        self.check_every = int(check_every)

    def table(self, steps):
        i = np.asarray(steps, dtype=np.double)[:, None]
        out = self.c[None, :] + self.l[None, :] * i
        if self.omega != 0.0:
            out = out + self.a[None, :] * np.cos(self.omega * i) + self.b[None, :] * np.sin(self.omega * i)
        return out

    @staticmethod
    def _design(steps, omega):
        i = np.asarray(steps, dtype=np.double)
        cols = [np.ones_like(i), i]
        if omega != 0.0:
            cols += [np.cos(omega * i), np.sin(omega * i)]
        return np.stack(cols, axis=1)

    @classmethod
    def _solve(cls, steps, X, omega):
        M = cls._design(steps, omega)
        coef, *_ = np.linalg.lstsq(M, X, rcond=None)
        return coef, X - M @ coef

    @classmethod
    def fit(cls, sample, Nt):
        near = list(range(min(cls.NEAR, Nt)))
        if len(near) < 6:
            return None
        far = sorted(set(int(round(f * (Nt - 1))) for f in (0.25, 0.5, 0.75, 1.0)) - set(near))
        steps = near + far
        import time
        t0 = time.perf_counter()
        X = np.stack([np.asarray(sample(i), dtype=np.double) for i in steps])        # (S, W)
        per_call = (time.perf_counter() - t0) / len(steps)
        check_every = int(min(4096, max(16, round(per_call / 50e-6))))
        scale = float(np.max(np.abs(X)))
        if scale == 0.0 or not np.all(np.isfinite(X)):
            return None
        tol = cls.RTOL * scale
        # second differences of the consecutive samples remove constant and ramp; what is left is the oscillation
        Xn = X[:len(near)]
        e = Xn[2:] - 2.0 * Xn[1:-1] + Xn[:-2]
        omega = 0.0
        if np.max(np.abs(e)) > tol:
            # a sinusoid obeys e[i+1] + e[i-1] = 2 cos(omega) e[i]
            den = float(np.sum(e[1:-1] ** 2))
            if den == 0.0:
                return None
            g = float(np.sum(e[1:-1] * (e[2:] + e[:-2]))) / den
            if not -2.0 < g < 2.0:
                return None
            omega = float(np.arccos(0.5 * g))
            # the frequency from ten neighbouring samples is good to ~1e-13; the far samples pin it (Gauss-Newton on omega,
            # a few steps, each bounded so that the phase at the far samples never slips by a cycle)
            i_all = np.asarray(steps, dtype=np.double)
            for _ in range(12):
                coef, r = cls._solve(steps, X, omega)
                if np.max(np.abs(r)) <= tol:
                    break
                dM = (-coef[2][None, :] * np.sin(omega * i_all)[:, None] + coef[3][None, :] * np.cos(omega * i_all)[:, None]) \
                    * i_all[:, None]
                # variable projection: the coefficients are re-fitted at every frequency, so only the part of the
                # derivative outside the span of the design columns moves the residual
                Q, _ = np.linalg.qr(cls._design(steps, omega))
                dM = dM - Q @ (Q.T @ dM)
                den = float(np.sum(dM * dM))
                if den == 0.0:
                    break
                step = float(np.sum(dM * r)) / den
                omega += float(np.clip(step, -0.25 / max(i_all[-1], 1.0), 0.25 / max(i_all[-1], 1.0)))
        coef, r = cls._solve(steps, X, omega)
        if np.max(np.abs(r)) > tol:
            return None
        W = X.shape[1]
        a, b = (coef[2], coef[3]) if omega != 0.0 else (np.zeros(W), np.zeros(W))
        l = np.where(np.abs(coef[1]) * max(Nt, 1) <= tol, 0.0, coef[1])
        return cls(coef[0], l, a, b, omega, check_every)


class SourceSpec:
    """Device-facing description of one input; ``chunk(i0, i1)`` yields the tables for steps [i0, i1)."""

    def __init__(self, kind, N, W, Nt, base=None, static=True, amp_fn=None, dense_fn=None, model=None):
        self.kind, self.N, self.W, self.Nt = kind, N, W, Nt
        self.model = model          # AmplitudeModel standing in for amp_fn between its checks, or None
        self.model_checks = 0       # how often the callable was evaluated to re-verify the model
        self.base = base            # (N,) for RANK1
        self.static = static        # True: a single table row is valid for every step
        self._amp_fn = amp_fn       # i -> (W,)
        self._dense_fn = dense_fn   # i -> (N, W)

    def is_zero(self):
        return self.kind == ZERO

    def amp_chunk(self, i0, i1):
        """(K, W) float64 amplitudes for RANK1; K == 1 when static."""
        steps = [i0] if self.static else range(i0, i1)
        if self.model is not None and not self.static:
            tab = self.model.table(steps)
            # re-verify against the callable at the multiples of check_every that fall into this chunk (and at its
            # first step); a miss retires the model for the rest of the run
            every = self.model.check_every
            checks = sorted(set([i0] + list(range(-(-i0 // every) * every, i1, every))))
            scale = max(float(np.max(np.abs(tab))), 1e-300)
            for i in checks:
                self.model_checks += 1
                if np.max(np.abs(np.asarray(self._amp_fn(i), dtype=np.double) - tab[i - i0])) > 10 * self.model.RTOL * scale:
                    self.model = None
                    break
            else:
                return np.ascontiguousarray(tab)
        return np.ascontiguousarray(np.stack([np.asarray(self._amp_fn(i), dtype=np.double) for i in steps]))

    def dense_chunk(self, i0, i1):
        """(K, N, W) float64 values for DENSE; K == 1 when static."""
        steps = [i0] if self.static else range(i0, i1)
        return np.ascontiguousarray(np.stack(
            [np.broadcast_to(np.asarray(self._dense_fn(i), dtype=np.double), (self.N, self.W)) for i in steps]))

    def as_dense(self):
        """The same input as per-step dense tables (for a callable that stopped being rank one)."""
        if self.kind == DENSE:
            return self
        fn = self._dense_fn
        if fn is None:
            base, amp_fn = self.base, self._amp_fn
            fn = (lambda i: base[:, None] * np.asarray(amp_fn(i), dtype=np.double)[None, :]) if self.kind == RANK1 \
                else (lambda i: np.zeros((self.N, self.W)))
        return SourceSpec(DENSE, self.N, self.W, self.Nt, static=self.static, dense_fn=fn)

    def value(self, i):
        """Dense (N, W) value at step i (host; used for stored currents and tests)."""
        if self.kind == ZERO:
            return np.zeros((self.N, self.W))
        if self.kind == RANK1:
            return self.base[:, None] * np.asarray(self._amp_fn(i), dtype=np.double)[None, :]
        return np.broadcast_to(np.asarray(self._dense_fn(i), dtype=np.double), (self.N, self.W))


def _rank_one_factor(a2, rtol=_RANK1_RTOL):
    """Try a2 (N, W) == base (N,) x amp (W,). Returns (base, amp) or None."""
    N, W = a2.shape
    flat = np.argmax(np.abs(a2))
    j0, w0 = divmod(int(flat), W)
    piv = a2[j0, w0]
    if piv == 0.0:
        return np.zeros(N), np.zeros(W)
    amp = a2[j0, :].copy()
    base = a2[:, w0] / piv
    if np.allclose(base[:, None] * amp[None, :], a2, rtol=rtol, atol=0.0):
        return base, amp
    return None


def _collapse_broadcast(arr):
    """(a, b, c) view of an array of at most three dimensions with every zero-stride axis (np.broadcast_to views, as
    a reference-style problem object stores its inputs: time_evolution.py:119-130) reduced to length one."""
    a3 = arr.reshape((1,) * (3 - arr.ndim) + arr.shape)
    index = tuple(slice(0, 1) if (a3.strides[k] == 0 and a3.shape[k] > 1) else slice(None) for k in range(3))
    return a3[index]


_CHECK_ELEMS = 1 << 24          # elements per block when a time-dependent array is verified to be rank one


def _rank_one_over_time(a3, base, j0, N, W, Nt):
    """Is a3 (N, W, Nt) == base[:, None, None] * (a3[j0] / base[j0]) everywhere? Checked in blocks of time steps, so
    that no copy of the whole array is ever made."""
    step = max(1, _CHECK_ELEMS // max(1, N * W))
    for i0 in range(0, Nt, step):
        blk = a3[:, :, i0:i0 + step]
        tab = blk[j0] / base[j0]
        if not np.allclose(base[:, None, None] * tab[None, :, :], blk, rtol=_RANK1_RTOL, atol=0.0):
            return False
    return True


def classify_source(x, N, W, Nt, zero_if_allclose=True):
    """
    Build a SourceSpec from a reference-style input: scalar / array broadcastable to (N, W, Nt) /
    callable i -> broadcastable to (N, W).  Zero detection follows the reference: a time-independent
    input with np.allclose(x, 0) is dropped (reference: time_evolution.py:509-519, quirk Q5).
    """
    if isinstance(x, RankOneSource):
        if x.base.size != N:
            raise ValueError("RankOneSource base has wrong length")
        return SourceSpec(RANK1, N, W, Nt, base=x.base, static=x.is_static(), amp_fn=x.amp)
    if callable(x):
        # generic callable: time dependent by definition (reference: time_evolution.py:334-336).
        first = np.broadcast_to(np.asarray(x(0), dtype=np.double), (N, W))
        rf = _rank_one_factor(first)
        if rf is not None and np.any(rf[0] != 0.0):
            base = rf[0]
            j0 = int(np.argmax(np.abs(base)))     # base[j0] == 1 by construction

            def amp_fn(i, _x=x, _base=base, _j0=j0):
                v = np.broadcast_to(np.asarray(_x(i), dtype=np.double), (N, W))
                amp = v[_j0, :] / _base[_j0]
                if not np.allclose(_base[:, None] * amp[None, :], v, rtol=_RANK1_RTOL, atol=0.0):
                    raise NotRankOne()
                return amp
            # probe a few steps: a callable whose structure changes over time is handled densely (and one that
            # changes later in the run makes the engine switch to the dense form from that chunk of steps on)
            try:
                for i in sorted(set([0, 1, Nt // 3, Nt // 2, Nt - 1])):
                    if 0 <= i < Nt:
                        amp_fn(i)
                import os
                use_model = Nt > 4 * AmplitudeModel.NEAR and os.environ.get("JJ_SOURCE_MODEL", "1") != "0"
                model = AmplitudeModel.fit(amp_fn, Nt) if use_model else None
                return SourceSpec(RANK1, N, W, Nt, base=base, static=False, amp_fn=amp_fn, dense_fn=x, model=model)
            except NotRankOne:
                pass
        return SourceSpec(DENSE, N, W, Nt, static=False, dense_fn=x)
    arr = np.asarray(x, dtype=np.double)
    np.broadcast_to(arr, (N, W, Nt))             # raises like the reference on bad shapes
    if arr.ndim > 3:
        raise ValueError("input must be broadcastable to (N, W, Nt)")
    a3 = _collapse_broadcast(arr)                # zero-stride axes (broadcast views) count as length one
    full = np.broadcast_to(a3, (N, W, Nt))
    timedep = a3.shape[-1] > 1                   # reference: _is_timedep; a constant stored as a broadcast view is constant
    if not timedep:
        s0 = full[:, :, 0]
        if zero_if_allclose and np.allclose(a3[:, :, 0], 0):      # same values as s0, without the broadcast copies
            return SourceSpec(ZERO, N, W, Nt)
        # exploit broadcast structure first (exact), then a numerical rank-one test
        if a3.shape[0] == 1:
            amp = np.broadcast_to(a3[0, :, 0], (W,)).copy()
            return SourceSpec(RANK1, N, W, Nt, base=np.ones(N), static=True, amp_fn=lambda i, a=amp: a)
        if a3.shape[1] == 1:
            base = np.ascontiguousarray(a3[:, 0, 0])
            one = np.ones(W)
            return SourceSpec(RANK1, N, W, Nt, base=base, static=True, amp_fn=lambda i, a=one: a)
        rf = _rank_one_factor(np.ascontiguousarray(s0))
        if rf is not None:
            return SourceSpec(RANK1, N, W, Nt, base=rf[0], static=True, amp_fn=lambda i, a=rf[1]: a)
        dense = np.ascontiguousarray(s0)
        return SourceSpec(DENSE, N, W, Nt, static=True, dense_fn=lambda i, d=dense: d)
    # time dependent array
    if a3.shape[0] == 1:
        tab = np.broadcast_to(a3[0], (W, Nt))
        return SourceSpec(RANK1, N, W, Nt, base=np.ones(N), static=False, amp_fn=lambda i, t=tab: t[:, i])
    if a3.shape[1] == 1 or W == 1:
        # (N, 1, Nt): rank one in (e, w) for every step with amp = 1, but base changes per step -> dense
        return SourceSpec(DENSE, N, W, Nt, static=False, dense_fn=lambda i, f=full: f[:, :, i])
    rf = _rank_one_factor(np.ascontiguousarray(full[:, :, 0]))
    if rf is not None and np.any(rf[0] != 0.0):
        base = rf[0]
        j0 = int(np.argmax(np.abs(base)))
        if _rank_one_over_time(full, base, j0, N, W, Nt):
            tab = full[j0, :, :] / base[j0]
            return SourceSpec(RANK1, N, W, Nt, base=base, static=False, amp_fn=lambda i, t=tab: t[:, i])
    return SourceSpec(DENSE, N, W, Nt, static=False, dense_fn=lambda i, f=full: f[:, :, i])


class NotRankOne(ValueError):
    """A callable input stopped being of the form base[e] * amp(i)[w]; the engine catches this and continues with the
    dense form of that input (SourceSpec.as_dense)."""

    def __init__(self):
        super().__init__("a callable input stopped being of the form base[e] * amp(i)[w] during the run")


def nonnegative_factors(spec):
    """For temperature: make base >= 0 and amp >= 0 (T = base * amp >= 0) or fall back to DENSE."""
    if spec.kind != RANK1:
        return spec
    if np.all(spec.base >= 0):
        return spec
    if np.all(spec.base <= 0):
        fn = spec._amp_fn
        return SourceSpec(RANK1, spec.N, spec.W, spec.Nt, base=-spec.base, static=spec.static,
                          amp_fn=lambda i, f=fn: -np.asarray(f(i)))
    base, fn = spec.base, spec._amp_fn
    return SourceSpec(DENSE, spec.N, spec.W, spec.Nt, static=spec.static,
                      dense_fn=lambda i, b=base, f=fn: b[:, None] * np.asarray(f(i))[None, :])
