"""
``TimeEvolutionResult``: the stored planes of a time evolution and the quantities derived from them
(the reference's container and getter names, time_evolution.py:597-1067).

Every derived quantity is evaluated for ALL selected time points in one batched operation - the planes are viewed as
one (N, W * K) matrix, so a getter is one sparse product or one multi-right-hand-side solve instead of the reference's
Python loop over time points - and is described by a row of ``_DERIVED``: which stored array it needs, the row count
of its output, and the batched formula. Beyond the reference: the running observables the device accumulated while
stepping (``get_mean_vortex_configuration``, ``get_dc_voltage``, ...) and the vortex configurations computed on the
device (``store_vortex_configuration``).
"""
import numpy as np

__all__ = ["TimeEvolutionResult"]

TWO_PI = 2.0 * np.pi


def _exceptions():
    from . import errors
    return errors


class _Batch:
    """The selected planes of the stored arrays as (N, W * K) matrices plus what the formulas need to know."""

    def __init__(self, result, steps):
        self.result, self.problem, self.circuit = result, result.problem, result.problem.circuit
        self.steps = steps
        self.W, self.K = result.get_problem_count(), len(steps)

    def flat(self, kind):
        return self.result._stored(kind, self.steps).reshape(-1, self.W * self.K)

    def input(self, key):
        """a per-step input at the selected steps as (N, W * K)"""
        per_step = [getattr(self.problem, "_" + key)(int(t)) for t in self.steps]
        return np.stack(per_step, axis=2).reshape(per_step[0].shape[0], -1) if per_step else np.zeros((0, 0))


def _vortices(b):
    A = b.circuit.get_cycle_matrix()
    return -(A @ np.round(b.flat("theta") / TWO_PI))


def _node_phase(b):
    c = b.circuit
    return c.Msq_solve(c.get_cut_matrix() @ b.flat("theta"))


def _potential(b):
    c = b.circuit
    return c.Msq_solve(c.get_cut_matrix() @ b.flat("voltage"))


def _cycle_current(b):
    c = b.circuit
    return c.Asq_solve(c.get_cycle_matrix() @ (b.flat("current") - b.input("Is")))


def _flux(b):
    c = b.circuit
    return b.input("f") + c.get_cycle_matrix() @ (c._L() @ b.flat("current")) / TWO_PI


# name -> (stored array it is derived from, rows of the result, batched formula, circuit predicate that makes it zero)
_DERIVED = {
    "phase": ("theta", "_Nn", _node_phase, None),
    "vortex_configuration": ("theta", "_Nf", _vortices, None),
    "josephson_energy": ("theta", "_Nj", lambda b: b.problem._icp(b.flat("theta")), None),
    "supercurrent": ("theta", "_Nj", lambda b: b.problem._cp(b.flat("theta")), None),
    "cycle_current": ("current", "_Nf", _cycle_current, None),
    "flux": ("current", "_Nf", _flux, None),
    "magnetic_energy": ("current", "_Nj", lambda b: 0.5 * (b.circuit._L() @ b.flat("current") ** 2),
                        lambda c: not c._has_inductance()),
    "potential": ("voltage", "_Nn", _potential, None),
    "capacitive_energy": ("voltage", "_Nj", lambda b: 0.5 * b.circuit._C()[:, None] * b.flat("voltage") ** 2,
                          lambda c: not c._has_capacitance()),
}

_MISSING = {"theta": ("ThetaNotStored", "theta"), "current": ("CurrentNotStored", "current"),
            "voltage": ("VoltageNotStored", "voltage")}


class TimeEvolutionResult:
    """
    theta, current, voltage: (Nj, W, Nt_s) arrays over the stored steps, or None when not stored.
    ``observed``: what the device accumulated beside the planes (filled by time_evolution()).
    """

    def __init__(self, problem, theta, current, voltage, observed=None):
        self.problem = problem
        shape = (problem.circuit._Nj(), problem.get_problem_count(), problem._Nt_s())
        for name, arr, wanted in (("theta", theta, problem.store_theta), ("current", current, problem.store_current),
                                  ("voltage", voltage, problem.store_voltage)):
            if wanted and np.shape(arr) != shape:
                raise ValueError(f"{name} must have shape {shape}; has shape {np.shape(arr)}")
            setattr(self, name, arr if wanted else None)
        stored = problem.store_time_steps
        # plane of every step among the stored ones (meaningful where stored)
        self.time_point_indices = np.cumsum(stored) - stored.astype(int)
        self.animation = None
        self.observed = observed or {}

    # ---- selection
    def _steps(self, select_time_points):
        steps = np.flatnonzero(self.problem._to_time_point_mask(select_time_points))
        if not np.all(self.problem.store_time_steps[steps]):
            raise _exceptions().DataAtTimepointNotStored("Queried a timepoint that is not stored during time evolution simulation.")
        return steps

    def _time_point_index(self, time_points):
        if time_points is None:
            time_points = self.problem.store_time_steps
        if not np.all(self.problem.store_time_steps[time_points]):
            raise _exceptions().DataAtTimepointNotStored("Queried a timepoint that is not stored during time evolution simulation.")
        return self.time_point_indices[time_points]

    def _stored(self, kind, steps, purpose=None):
        arr = getattr(self, kind)
        if arr is None:
            exc, label = _MISSING[kind]
            what = f"Cannot compute {purpose}; requires {label} to be stored in TimeEvolutionConfig" if purpose else \
                f"Cannot query {label}; quantity is not stored during time evolution simulation."
            raise getattr(_exceptions(), exc)(what)
        return arr[:, :, self.time_point_indices[steps]]

    def _th(self, time_point):
        return self._stored("theta", np.atleast_1d(time_point))[:, :, 0] if np.ndim(time_point) == 0 else \
            self._stored("theta", time_point)

    def _I(self, time_point):
        return self._stored("current", np.atleast_1d(time_point))[:, :, 0] if np.ndim(time_point) == 0 else \
            self._stored("current", time_point)

    def _V(self, time_point):
        return self._stored("voltage", np.atleast_1d(time_point))[:, :, 0] if np.ndim(time_point) == 0 else \
            self._stored("voltage", time_point)

    # ---- the reference's accessors
    def get_problem_count(self):
        return self.problem.get_problem_count()

    def get_circuit(self):
        return self.problem.get_circuit()

    def select_static_configuration(self, prob_nr, time_step):
        raise NotImplementedError("static configurations are outside the time-evolution hot path "
                                  "(reference: static_problem.py:815; SURVEY.md section 2 row C10)")

    def get_theta(self, select_time_points=None):
        return np.array(self._stored("theta", self._steps(select_time_points)), dtype=np.double)

    def get_current(self, select_time_points=None):
        return np.array(self._stored("current", self._steps(select_time_points)), dtype=np.double)

    def get_voltage(self, select_time_points=None):
        return np.array(self._stored("voltage", self._steps(select_time_points)), dtype=np.double)

    def _derive(self, name, select_time_points):
        source, rows, formula, zero_when = _DERIVED[name]
        steps = self._steps(select_time_points)
        c = self.get_circuit()
        N, W, K = getattr(c, rows)(), self.get_problem_count(), len(steps)
        self._stored(source, steps[:0], purpose=name.replace("_", " "))     # raises when the source is not stored
        if K == 0 or (zero_when is not None and zero_when(c)):
            return np.zeros((N, W, K))
        return np.asarray(formula(_Batch(self, steps)), dtype=np.double).reshape(N, W, K)

    def get_vortex_configuration(self, select_time_points=None):
        """n = -A round(theta / 2 pi), (Nf, W, K) int. Taken from the planes the device computed when the problem
        asked for them (store_vortex_configuration: works without stored phases), otherwise derived from theta."""
        planes = self.observed.get("n_planes")
        if planes is not None:
            steps = self._steps(select_time_points)
            idx = self.time_point_indices[steps]
            if not (idx.size == planes.shape[0] and np.array_equal(idx, np.arange(idx.size))):
                planes = planes[idx]
            # (widened plane by plane in memory order, then viewed in the reference's (Nf, W, K) layout like the phases)
            return np.moveaxis(planes.astype(int), 0, 2)
        return self._derive("vortex_configuration", select_time_points).astype(int)

    def get_energy(self, select_time_points=None):
        return self.get_josephson_energy(select_time_points) + self.get_magnetic_energy(select_time_points) + \
            self.get_capacitive_energy(select_time_points)

    # ---- running observables accumulated on the device while stepping (observe_interval)
    def _observed(self, key):
        if not self.observed.get("interval"):
            raise ValueError("no running observables: construct the problem with observe_interval=k")
        return self.observed[key]

    def get_observation_count(self):
        return int(self._observed("count"))

    def get_observed_steps(self):
        first, k = self._observed("first"), self._observed("interval")
        return first + k * np.arange(self.get_observation_count())

    def get_vortex_sum(self):
        """(Nf, W) int: sum over the observed steps of the vortex configuration (exact)"""
        return self._observed("nsum").astype(int)

    def get_mean_vortex_configuration(self):
        """(Nf, W): time-averaged vortex configuration over the observed steps"""
        return self.get_vortex_sum() / max(1, self.get_observation_count())

    def get_dc_voltage(self):
        """(Nj, W): time-averaged junction voltage between the first and the latest observation,
        (theta_latest - theta_first) / elapsed time: the sum of the per-step voltages (theta(t) - theta(t-1)) / dt
        telescopes, so two phase marks on the device replace a stored voltage plane per step"""
        count = self.get_observation_count()
        if count < 2:
            raise ValueError("the DC voltage needs at least two observations")
        span = (count - 1) * self._observed("interval") * self.problem._dt()
        return (self._observed("theta_latest") - self._observed("theta_first")) / span

    def plot(self, *args, **kwargs):
        raise NotImplementedError("visualisation is outside the time-evolution hot path "
                                  "(reference: circuit_visualize.py; SURVEY.md section 2 row C13)")

    animate = plot

    def __str__(self):
        parts = [f"{label}{arr.shape}" for label, arr in (("th", self.theta), ("I", self.current), ("V", self.voltage))
                 if arr is not None]
        return f"time evolution configuration: ({', '.join(parts)})\nproblem: {self.problem}\ncircuit: {self.get_circuit()}"


def _install_derived(cls):
    for name in _DERIVED:
        if not hasattr(cls, "get_" + name):
            setattr(cls, "get_" + name, (lambda n: lambda self, select_time_points=None: self._derive(n, select_time_points))(name))


_install_derived(TimeEvolutionResult)
