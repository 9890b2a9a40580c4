"""
Current-phase relations. (reference: static_problem.py:34-85)

The reference lets ``func(Ic, theta)`` be any numpy ufunc. The device path evaluates
``Ic * g(theta)`` with ``g`` a 2 pi-periodic trigonometric polynomial; ``harmonics()`` measures the
coefficients of ``g`` from the callable and verifies the fit, raising if the relation is not of that
form (there is no CPU fallback).
"""
import numpy as np

__all__ = ["CurrentPhaseRelation", "DefaultCPR"]

_MAX_HARMONICS = 16


class CurrentPhaseRelation:
    """
    Current-phase relation Icp(Ic, theta) with derivative and integral over theta.
    (reference: static_problem.py:34-75)
    """

    def __init__(self, func, d_func, i_func):
        self.func = func
        self.d_func = d_func
        self.i_func = i_func

    def eval(self, Ic, theta):
        return self.func(Ic, theta)

    def d_eval(self, Ic, theta):
        return self.d_func(Ic, theta)

    def i_eval(self, Ic, theta):
        return self.i_func(Ic, theta)


class DefaultCPR(CurrentPhaseRelation):
    """Icp = Ic sin(theta). (reference: static_problem.py:77-85)"""

    def __init__(self):
        super().__init__(lambda Ic, th: Ic * np.sin(th),
                         lambda Ic, th: Ic * np.cos(th),
                         lambda Ic, th: Ic * (1.0 - np.cos(th)))


def harmonics(cpr, tol=1e-11):
    """
    Coefficients (a_0..a_M, b_1..b_M) with func(Ic, th) = Ic * (sum_m a_m cos(m th) + b_m sin(m th)).

    Returns (a, b) as float64 arrays of equal length M+1 (b[0] = 0). A DefaultCPR-equivalent relation
    returns a = [0, 0], b = [0, 1]. Raises ValueError when the callable is not proportional to Ic or
    not a trigonometric polynomial of degree <= 16 (checked at random phases).
    """
    if isinstance(cpr, DefaultCPR) or type(cpr).__name__ == "DefaultCPR":
        return np.zeros(2), np.array([0.0, 1.0])
    K = 4 * _MAX_HARMONICS
    th = 2.0 * np.pi * np.arange(K) / K
    g = np.asarray(cpr.eval(np.ones(K), th), dtype=np.double)
    if g.shape != (K,) or not np.all(np.isfinite(g)):
        raise ValueError("current-phase relation must be a finite numpy ufunc of (Ic, theta)")
    F = np.fft.rfft(g) / K
    a = 2.0 * F.real
    b = -2.0 * F.imag
    a[0] = F[0].real
    a, b = a[:_MAX_HARMONICS + 1].copy(), b[:_MAX_HARMONICS + 1].copy()
    scale = max(np.max(np.abs(g)), 1e-300)
    a[np.abs(a) < 1e-14 * scale] = 0.0
    b[np.abs(b) < 1e-14 * scale] = 0.0
    b[0] = 0.0
    M = max([m for m in range(_MAX_HARMONICS + 1) if a[m] != 0.0 or b[m] != 0.0] + [1])
    a, b = a[:M + 1], b[:M + 1]
    # verify: periodicity, trigonometric-polynomial form, linearity in Ic
    rng = np.random.RandomState(12345)
    tt = rng.uniform(-40.0, 40.0, 257)
    m = np.arange(M + 1)[:, None]
    fit = (a[:, None] * np.cos(m * tt) + b[:, None] * np.sin(m * tt)).sum(axis=0)
    ic = rng.uniform(0.1, 3.0, 257)
    got = np.asarray(cpr.eval(ic, tt), dtype=np.double)
    if not np.allclose(got, ic * fit, rtol=0, atol=tol * scale * 3.0):
        raise ValueError("current-phase relation is not of the form Ic * (trigonometric polynomial of "
                         "degree <= %d in theta); the device path cannot evaluate it" % _MAX_HARMONICS)
    return a, b
