"""
Time evolution on Josephson circuits: the reference's public entry point for its hot path, backed by the B200 engine.

``TimeEvolutionProblem`` takes the reference's constructor arguments and offers its getters
(reference: time_evolution.py:20-408), so scripts written against the reference run unchanged; what happens inside is
this package's own: inputs are described by one table (``_INPUTS``), the initial conditions are held lazily (an
all-zero state is never materialised, copied or uploaded), ``compute()`` plans which steps the device must keep
(``StorePlan``) and hands the stepping loop (reference: time_evolution.py:461-582) to
``engine.device_time_evolution_core``, which drives hand-written sm_100a kernels through the C ABI of
include/jjstep.h. There is no CPU fallback: without the CUDA library or a GPU, ``compute()`` raises.

Optional keyword-only arguments beyond the reference's (defaults change nothing):
  noise_seed      int, seed of the counter-based Philox generator of the thermal noise (None: a fresh seed per
                  compute(), drawn from numpy's global generator, so ``np.random.seed`` still makes a script repeatable)
  noise_replay    (Nt, Nj, W) array or callable i -> (Nj, W) of standard normal draws to use instead of the device
                  generator (parity tests against the reference's MT19937 stream)
  devices         CUDA device ordinals to shard the problem axis over from host threads
  observe_interval, observe_first
                  running observables accumulated ON THE DEVICE while stepping (no theta plane is stored or copied):
                  steps observe_first + m * observe_interval add their vortex configuration to a per-(face, problem)
                  sum and mark their phases; the result offers the time-averaged vortex configuration and the DC
                  junction voltages over the observed window (``TimeEvolutionResult.get_mean_vortex_configuration``,
                  ``get_dc_voltage``)
  store_vortex_configuration
                  keep n = -A round(theta / 2 pi) of every stored step, computed on the device from the stored phase
                  planes and copied out as integers; with store_theta=False the phases never leave the GPU
"""
from __future__ import annotations

import numpy as np

from .current_phase_relation import DefaultCPR
from .errors import ThetaNotStored, CurrentNotStored, VoltageNotStored, DataAtTimepointNotStored
from .josephson_circuit import Circuit

__all__ = ["TimeEvolutionProblem", "TimeEvolutionResult", "ThetaNotStored", "CurrentNotStored",
           "VoltageNotStored", "DataAtTimepointNotStored", "time_evolution", "AnnealingProblem",
           "AnnealedConfiguration", "StorePlan"]


# the four per-step inputs: attribute, key of the device-side classifier, rows ("f": per face, else per junction),
# name of the time-dependence flag (frozen at construction like the reference's, time_evolution.py:117-131)
_INPUTS = (("external_flux", "f", "_Nf", "_f_is_timedep"),
           ("current_sources", "Is", "_Nj", "_Is_is_timedep"),
           ("voltage_sources", "Vs", "_Nj", "_Vs_is_timedep"),
           ("temperature", "T", "_Nj", "_T_is_timedep"))


def _varies_in_time(x):
    """callables and arrays with a time axis longer than one (reference: time_evolution.py:334-340)"""
    if callable(x):
        return True
    shape = np.shape(x)
    return len(shape) > 0 and shape[-1] > 1


def _problem_axis(x):
    """length of the problem axis an input brings along (1 when it has none)"""
    if callable(x):
        declared = getattr(x, "problem_count", None)
        if declared is not None:
            return int(declared)
        x = x(0)
    shape = np.shape(x)
    return shape[1] if len(shape) > 1 else 1


class TimeEvolutionProblem:
    """
    W time-evolution problems on one circuit, integrated in lockstep on the GPU.

    Parameters (the reference's, time_evolution.py:20-149)
    ----------
    circuit : Circuit
    time_step=0.05, time_step_count=1000 : dt, Nt
    current_phase_relation=DefaultCPR()
    external_flux, current_sources, voltage_sources, temperature = 0.0 :
        array broadcastable to (Nf or Nj, W, Nt), or a callable step -> array broadcastable to (Nf or Nj, W)
    store_time_steps=None : indices in range(Nt), mask of shape (Nt,), or None (all steps)
    store_theta, store_voltage, store_current = True
    config_at_minus_1, config_at_minus_2 = None : (Nj, W) arrays or objects with get_theta(); None = at rest
    stencil_width=3 : only 3 (4 and 5 are inconsistent in the reference)
    """

    def __init__(self, circuit: Circuit, time_step=0.05, time_step_count=1000,
                 current_phase_relation=DefaultCPR(),
                 external_flux=0.0, current_sources=0.0,
                 voltage_sources=0.0, temperature=0.0,
                 store_time_steps=None, store_theta=True, store_voltage=True, store_current=True,
                 config_at_minus_1: np.ndarray = None,
                 config_at_minus_2: np.ndarray = None,
                 config_at_minus_3: np.ndarray = None,
                 config_at_minus_4: np.ndarray = None,
                 stencil_width=3, *, noise_seed=None, noise_replay=None, devices=None,
                 observe_interval=None, observe_first=0, store_vortex_configuration=False):
        self.circuit, self.time_step, self.time_step_count = circuit, time_step, time_step_count
        self.current_phase_relation = current_phase_relation
        given = dict(external_flux=external_flux, current_sources=current_sources,
                     voltage_sources=voltage_sources, temperature=temperature)
        self.problem_count = max(_problem_axis(v) for v in given.values())
        W, Nt = self.problem_count, time_step_count
        # what the device-side classifier reads (sources.classify_source): the inputs as given, and the objects kept
        # under the public attribute names, so that an attribute replaced later (the reference's annealing loop
        # assigns prob.temperature, time_evolution.py:1166) is noticed by identity
        self._raw_sources, self._kept_sources = {}, {}
        for attr, key, rows, flag in _INPUTS:
            value = given[attr]
            setattr(self, flag, _varies_in_time(value))
            held = value if callable(value) else np.broadcast_to(np.array(value), (getattr(circuit, rows)(), W, Nt))
            setattr(self, attr, held)
            self._raw_sources[key], self._kept_sources[key] = value, held

        self.store_theta, self.store_voltage, self.store_current = store_theta, store_voltage, store_current
        self.store_vortex_configuration = bool(store_vortex_configuration)
        self.observe_interval = int(observe_interval) if observe_interval else 0
        self.observe_first = int(observe_first)
        if self.observe_interval < 0 or self.observe_first < 0:
            raise ValueError("observe_interval and observe_first must not be negative")
        self.store_time_steps = self._to_time_point_mask(store_time_steps)
        stores_planes = store_theta or store_voltage or store_current or self.store_vortex_configuration
        if not (stores_planes or self.observe_interval):
            raise ValueError("No output is stored")
        if not self.store_time_steps.any() and not self.observe_interval:
            raise ValueError("No output is stored")
        self.stencil_width = stencil_width
        self.stencil = self._get_stencil(stencil_width)
        self._m1 = self._as_state(config_at_minus_1)
        self._m2 = self._as_state(config_at_minus_2)
        self.noise_seed, self.noise_replay, self.devices = noise_seed, noise_replay, devices

    # ---- initial conditions: None stands for "like the step after" (theta(-2) := theta(-1) := rest) and is only
    # turned into an array when somebody reads the attribute
    def _as_state(self, cfg):
        if cfg is None:
            return None
        if hasattr(cfg, "get_theta"):
            cfg = cfg.get_theta()
        return np.asarray(cfg).reshape(self.circuit._Nj(), self.problem_count)

    def starts_at_rest_with_zero_phases(self):
        """True when neither initial condition was given: the device state is simply cleared"""
        return self._m1 is None and self._m2 is None

    def initial_phases(self):
        """(theta(-1), theta(-2)) for the device upload, without materialising the copy that stands for an omitted
        theta(-2) (a problem that starts at rest)"""
        m1 = self.config_at_minus_1
        return m1, (m1 if self._m2 is None else self._m2)

    @property
    def config_at_minus_1(self):
        if self._m1 is None:
            self._m1 = np.zeros((self.circuit._Nj(), self.problem_count))
        return self._m1

    @config_at_minus_1.setter
    def config_at_minus_1(self, cfg):
        self._m1 = self._as_state(cfg)

    @property
    def config_at_minus_2(self):
        if self._m2 is None:
            self._m2 = self.config_at_minus_1.copy()
        return self._m2

    @config_at_minus_2.setter
    def config_at_minus_2(self, cfg):
        self._m2 = self._as_state(cfg)

    # ---- the reference's getters (time_evolution.py:151-301)
    def get_static_problem(self, vortex_configuration, problem_nr=0, time_step=0):
        raise NotImplementedError("static problems are outside the time-evolution hot path "
                                  "(reference: static_problem.py; SURVEY.md section 2 row C9)")

    def get_phase_zone(self):
        return 0

    def get_net_sourced_current(self, time_step):
        return 0.5 * np.abs(self.get_node_current_sources(time_step)).sum(axis=0)

    def get_node_current_sources(self, time_step):
        return self.circuit.get_cut_matrix() @ self._Is(time_step)

    def get_time(self):
        return self.time_step * np.arange(self.time_step_count, dtype=np.double)

    def compute(self) -> "TimeEvolutionResult":
        """Integrate on the GPU (reference: time_evolution.py:303-307)."""
        return time_evolution(self)

    def __str__(self):
        lines = [f"time evolution problem: ", f"\ttime: {self.time_step_count} steps of {self.time_step}"]
        lines += [f"\t{label}: {getattr(self, attr)}" for label, attr in
                  (("current sources", "current_sources"), ("voltage sources", "voltage_sources"),
                   ("external_flux", "external_flux"), ("temperature", "temperature"),
                   ("current-phase relation", "current_phase_relation"))]
        return "\n".join(lines)

    # ---- internals under the reference's names (the device core and the result getters are duck-typed on them)
    def _Nt(self):
        return self.time_step_count

    def _Nt_s(self):
        return int(np.count_nonzero(self.store_time_steps))

    def _dt(self):
        return self.time_step

    def _input_at(self, attr, rows, step):
        x = getattr(self, attr)
        if callable(x):
            return np.broadcast_to(x(step), (getattr(self.circuit, rows)(), self.problem_count))
        return x[:, :, step]

    def _cp(self, theta):
        return self.current_phase_relation.eval(self.circuit._Ic()[:, None], theta)

    def _dcp(self, theta):
        return self.current_phase_relation.d_eval(self.circuit._Ic()[:, None], theta)

    def _icp(self, theta):
        return self.current_phase_relation.i_eval(self.circuit._Ic()[:, None], theta)

    def _to_time_point_mask(self, time_points):
        """None (the stored steps; every step while the problem is being constructed), a boolean mask over the steps,
        or anything that indexes an array of Nt steps -> boolean mask of shape (Nt,)"""
        if time_points is None:
            stored = getattr(self, "store_time_steps", None)
            return np.ones(self.time_step_count, dtype=bool) if stored is None else stored
        mask = np.zeros(self.time_step_count, dtype=bool)
        sel = time_points if isinstance(time_points, slice) else np.array(time_points)
        if not isinstance(sel, slice) and sel.dtype == bool:
            return sel
        try:
            mask[sel] = True
        except Exception:
            raise ValueError("Invalid store_time_steps; must be None, mask, slice or index array")
        return mask

    @staticmethod
    def _get_stencil(width: int):
        if width == 3:
            return (1.0, -1.0, 0.0), (1.0, -2.0, 1.0)
        if width in (4, 5):
            raise NotImplementedError(
                "stencil_width 4 and 5 are not provided: in the reference they use a system matrix that is "
                "inconsistent with the stencil (flux quantisation violated at 3e-3, SURVEY.md quirk Q1)")
        raise ValueError(f"stencil width must be 3, 4 or 5 (equals {width})")


def _install_accessors(cls):
    # get_<attribute>() for the plain attributes, and _f / _Is / _Vs / _T (step) -> (rows, W) slices
    for name in ("circuit", "problem_count", "time_step", "time_step_count", "current_phase_relation", "external_flux",
                 "current_sources", "voltage_sources", "temperature", "store_time_steps", "store_theta", "store_voltage",
                 "store_current"):
        setattr(cls, "get_" + name, (lambda attr: lambda self: getattr(self, attr))(name))
    for attr, key, rows, _ in _INPUTS:
        setattr(cls, "_" + key, (lambda a, r: lambda self, step: self._input_at(a, r, step))(attr, rows))


_install_accessors(TimeEvolutionProblem)


class StorePlan:
    """Which steps the device has to keep so that the requested outputs can be formed, and where the requested ones
    sit among the kept planes (reference: time_evolution.py:422-458 decides the same with index lists).

    The voltage of a stored step t is a backward difference, V(t) = (theta(t) - theta(t-1)) / dt (+ L (I(t) - I(t-1)) / dt
    with inductance): step t - 1 is kept as a helper plane next to t. The device returns the kept planes behind the two
    initial conditions [theta(-2), theta(-1), kept...], so the plane before a kept step t is always at position - 1:
    its helper when t > 0, theta(-1) when t = 0."""

    def __init__(self, problem):
        store = problem.store_time_steps
        want_V = bool(problem.store_voltage) and store.any()
        has_L = problem.circuit._has_inductance()
        helper = np.zeros_like(store)
        if want_V:
            helper[:-1] = store[1:]
        none = np.zeros_like(store)
        want_n = getattr(problem, "store_vortex_configuration", False)
        self.requested = store
        self.theta_public = store if problem.store_theta else none
        self.current_public = store if problem.store_current else none
        self.theta_mask = (store | helper) if want_V else (store if (problem.store_theta or want_n) else none)
        # (the reference also keeps the currents of the stored steps whenever voltages are stored; they are only read
        # when there is inductance, so they are not kept - nor computed, nor copied - otherwise)
        self.current_mask = (store | helper) if (want_V and has_L) else (store if problem.store_current else none)
        self.want_V, self.has_L = want_V, has_L
        # theta leaves the device only when the host needs it (as output or for the voltage)
        self.fetch_theta = bool(problem.store_theta) or want_V

    @staticmethod
    def positions(kept_mask, wanted_mask):
        """plane positions (behind the two initial conditions) of the wanted steps among the kept ones"""
        return (np.cumsum(kept_mask) + 1)[wanted_mask & kept_mask]


def time_evolution(problem: TimeEvolutionProblem, core=None):
    """
    Run the stepping loop on the device and assemble the requested outputs. ``core`` (default: the GPU engine) has
    the signature of the reference's time_evolution_core: (problem, theta mask, current mask) -> two
    (Nj, W, kept + 2) arrays whose first two planes are the initial conditions.
    """
    plan = StorePlan(problem)
    extras = None
    wants_extras = bool(getattr(problem, "observe_interval", 0)) or bool(getattr(problem, "store_vortex_configuration", False))
    if core is None or getattr(core, "accepts_extras", False):
        extras = dict(interval=getattr(problem, "observe_interval", 0), first=getattr(problem, "observe_first", 0),
                      vortex_planes=bool(getattr(problem, "store_vortex_configuration", False)),
                      fetch_theta=plan.fetch_theta, wanted=plan.requested)
    if core is None:
        from .engine import device_time_evolution_core
        # the initial-condition planes are only read by the voltage difference of step 0
        th, I = device_time_evolution_core(problem, plan.theta_mask, plan.current_mask,
                                           initial_planes=plan.want_V, extras=extras)
    elif extras is not None:
        extras["initial_planes"] = plan.want_V
        th, I = core(problem, plan.theta_mask, plan.current_mask, extras=extras)
    else:
        if wants_extras:
            raise NotImplementedError("running observables and device-side vortex planes need the device core; the core "
                                      "passed to time_evolution() does not produce them")
        th, I = core(problem, plan.theta_mask, plan.current_mask)

    dt = problem._dt()
    lead = extras.get("lead", 2) if extras is not None else 2       # the device core drops the two initial planes when unread
    V = None
    if plan.want_V:
        at = StorePlan.positions(plan.theta_mask, plan.requested)
        V = (th[:, :, at] - th[:, :, at - 1]) / dt
        if plan.has_L:
            ai = StorePlan.positions(plan.current_mask, plan.requested)
            dI = (I[:, :, ai] - I[:, :, ai - 1]) / dt
            V += (problem.circuit.get_inductance() @ dI.reshape(dI.shape[0], -1)).reshape(V.shape)
    theta = current = None
    if problem.store_theta:
        theta = _planes(th, plan.theta_mask, plan.theta_public, lead)
    if problem.store_current:
        current = _planes(I, plan.current_mask, plan.current_public, lead)
    return TimeEvolutionResult(problem, theta, current, V if problem.store_voltage else None, observed=extras)


def _planes(arr, kept_mask, wanted_mask, lead=2):
    """the wanted steps out of the kept planes (behind `lead` initial planes); a plain slice (no copy) when every kept
    plane is wanted"""
    if np.array_equal(kept_mask, wanted_mask):
        return arr[:, :, lead:]
    return arr[:, :, StorePlan.positions(kept_mask, wanted_mask) - (2 - lead)]


from .result import TimeEvolutionResult                               # noqa: E402
from .annealing import AnnealingProblem, AnnealedConfiguration         # noqa: E402
