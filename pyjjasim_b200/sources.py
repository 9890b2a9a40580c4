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

__all__ = ["RankOneSource", "SourceSpec", "classify_source"]

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


class SourceSpec:
    """Device-facing description of one input; ``chunk(i0, i1)`` yields the tables for steps [i0, i1)."""

    def __init__(self, kind, N, W, Nt, base=None, static=True, amp_fn=None, dense_fn=None):
        self.kind, self.N, self.W, self.Nt = kind, N, W, Nt
        self.base = base            # (N,) for RANK1
        self.static = static        # True: a single table row is valid for every step
        self._amp_fn = amp_fn       # i -> (W,)
        self._dense_fn = dense_fn   # i -> (N, W)

    def is_zero(self):
        return self.kind == ZERO

    def amp_chunk(self, i0, i1):
        """(K, W) float64 amplitudes for RANK1; K == 1 when static."""
        steps = [i0] if self.static else range(i0, i1)
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
                return SourceSpec(RANK1, N, W, Nt, base=base, static=False, amp_fn=amp_fn, dense_fn=x)
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
