"""
Time evolution on Josephson circuits - the reference's public API for its hot path, backed by the
B200 device engine.

``TimeEvolutionProblem`` / ``TimeEvolutionResult`` mirror the reference's classes
(reference: time_evolution.py:20-408 and :597-1067): same constructor arguments, same getters,
same exceptions, same array shapes, so scripts written against the reference run unchanged.
What differs is below ``compute()``: the Python loop over numpy arrays
(reference: time_evolution.py:461-582) is replaced by ``engine.device_time_evolution_core`` which
drives hand-written sm_100a kernels through the C ABI in include/jjstep.h. There is no CPU
fallback: without the CUDA library or a GPU, ``compute()`` raises.

Extra, optional keyword arguments (defaults change nothing):
  noise_seed    : int, seed of the counter-based Philox generator used for thermal noise.
  noise_replay  : (Nt, Nj, W) array or callable i -> (Nj, W) of standard normal draws to use instead
                  of the device generator (parity testing against the reference's MT19937 stream).
  devices       : list of CUDA device ordinals to shard the problem axis over (default: current device).
"""
from __future__ import annotations

import numpy as np

from .current_phase_relation import DefaultCPR
from .josephson_circuit import Circuit

__all__ = ["TimeEvolutionProblem", "TimeEvolutionResult", "ThetaNotStored", "CurrentNotStored",
           "VoltageNotStored", "DataAtTimepointNotStored", "time_evolution", "AnnealingProblem",
           "AnnealedConfiguration"]


class ThetaNotStored(Exception):
    pass


class CurrentNotStored(Exception):
    pass


class VoltageNotStored(Exception):
    pass


class DataAtTimepointNotStored(Exception):
    pass


class TimeEvolutionProblem:
    """
    Define multiple time evolution problems with varying parameters in a Josephson circuit.
    All W problems are integrated in lockstep on the GPU. (reference: time_evolution.py:20-149)

    Parameters
    ----------
    circuit : Circuit
    time_step=0.05 : dt
    time_step_count=1000 : Nt
    current_phase_relation=DefaultCPR()
    external_flux=0.0 : array broadcastable to (Nf, W, Nt), or f(i) -> broadcastable to (Nf, W)
    current_sources=0.0 : array broadcastable to (Nj, W, Nt), or Is(i) -> broadcastable to (Nj, W)
    voltage_sources=0.0 : array broadcastable to (Nj, W, Nt), or Vs(i) -> ...
    temperature=0.0 : array broadcastable to (Nj, W, Nt), or T(i) -> ...
    store_time_steps=None : array in range(Nt), mask of shape (Nt,) or None (all)
    store_theta, store_voltage, store_current = True
    config_at_minus_1, config_at_minus_2 = None : (Nj, W) arrays or objects with get_theta()
    stencil_width=3 : only 3 is supported (4 and 5 are inconsistent in the reference)
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
                 stencil_width=3, *, noise_seed=None, noise_replay=None, devices=None):
        self.circuit = circuit
        self.time_step = time_step
        self.time_step_count = time_step_count
        self.current_phase_relation = current_phase_relation

        def get_prob_cnt(x):
            if hasattr(x, "problem_count") and hasattr(x, "__call__"):
                return x.problem_count
            s = np.array(x(0) if hasattr(x, "__call__") else x).shape
            return s[1] if len(s) > 1 else 1

        self.problem_count = max(get_prob_cnt(external_flux), get_prob_cnt(current_sources),
                                 get_prob_cnt(voltage_sources), get_prob_cnt(temperature))
        Nj, Nf, W, Nt = circuit._Nj(), circuit._Nf(), self.problem_count, self.time_step_count

        def keep(x, N):
            return x if hasattr(x, "__call__") else np.broadcast_to(np.array(x), (N, W, Nt))

        self._f_is_timedep = self._is_timedep(external_flux)
        self._Is_is_timedep = self._is_timedep(current_sources)
        self._Vs_is_timedep = self._is_timedep(voltage_sources)
        self._T_is_timedep = self._is_timedep(temperature)
        # raw inputs are kept for the device-side classifier (sources.classify_source)
        self._raw_sources = dict(f=external_flux, Is=current_sources, Vs=voltage_sources, T=temperature)
        self.external_flux = keep(external_flux, Nf)
        self.current_sources = keep(current_sources, Nj)
        self.voltage_sources = keep(voltage_sources, Nj)
        self.temperature = keep(temperature, Nj)
        self._kept_sources = dict(f=self.external_flux, Is=self.current_sources, Vs=self.voltage_sources,
                                  T=self.temperature)

        self.store_time_steps = np.ones(self._Nt(), dtype=bool)
        self.store_time_steps = self._to_time_point_mask(store_time_steps)
        self.store_theta = store_theta
        self.store_voltage = store_voltage
        self.store_current = store_current
        if not (self.store_theta or self.store_voltage or self.store_current):
            raise ValueError("No output is stored")
        if np.sum(self.store_time_steps) == 0:
            raise ValueError("No output is stored")
        self.stencil_width = stencil_width
        self.stencil = self._get_stencil(self.stencil_width)

        self.config_at_minus_1 = self._get_config(config_at_minus_1, np.zeros((Nj, W), dtype=np.double), (Nj, W))
        self.config_at_minus_2 = self._get_config(config_at_minus_2, self.config_at_minus_1, (Nj, W))
        self.noise_seed = noise_seed
        self.noise_replay = noise_replay
        self.devices = devices

    # --- getters (reference: time_evolution.py:151-301) ---------------------------------
    def get_static_problem(self, vortex_configuration, problem_nr=0, time_step=0):
        raise NotImplementedError("static problems are outside the time-evolution hot path "
                                  "(reference: static_problem.py; SURVEY.md section 2 row C9)")

    def get_problem_count(self):
        return self.problem_count

    def get_circuit(self) -> Circuit:
        return self.circuit

    def get_time_step(self):
        return self.time_step

    def get_time_step_count(self):
        return self.time_step_count

    def get_current_phase_relation(self):
        return self.current_phase_relation

    def get_phase_zone(self):
        return 0

    def get_external_flux(self):
        return self.external_flux

    def get_current_sources(self):
        return self.current_sources

    def get_net_sourced_current(self, time_step):
        M = self.get_circuit().get_cut_matrix()
        return 0.5 * np.sum(np.abs((M @ self._Is(time_step))), axis=0)

    def get_node_current_sources(self, time_step):
        M = self.get_circuit().get_cut_matrix()
        return M @ self._Is(time_step)

    def get_voltage_sources(self):
        return self.voltage_sources

    def get_temperature(self):
        return self.temperature

    def get_store_time_steps(self):
        return self.store_time_steps

    def get_store_theta(self):
        return self.store_theta

    def get_store_voltage(self):
        return self.store_voltage

    def get_store_current(self):
        return self.store_current

    def get_time(self):
        return np.arange(self._Nt(), dtype=np.double) * self._dt()

    def compute(self) -> "TimeEvolutionResult":
        """Compute the time evolution on the GPU. (reference: time_evolution.py:303-307)"""
        return time_evolution(self)

    def __str__(self):
        return "time evolution problem: " + \
               "\n\ttime: " + self.time_step_count.__str__() + " steps of " + self.time_step.__str__() + \
               "\n\tcurrent sources: " + self.current_sources.__str__() + \
               "\n\tvoltage sources: " + self.voltage_sources.__str__() + \
               "\n\texternal_flux: " + self.external_flux.__str__() + \
               "\n\ttemperature: " + self.temperature.__str__() + \
               "\n\tcurrent-phase relation: " + self.current_phase_relation.__str__()

    # --- internals with the reference's names ---------------------------------------------
    def _Nt(self):
        return self.time_step_count

    def _Nt_s(self):
        return int(np.sum(self.store_time_steps))

    def _dt(self):
        return self.time_step

    @staticmethod
    def _get_config(config_cur, config_prev, shape):
        config_cur = config_prev.copy() if config_cur is None else config_cur
        if hasattr(config_cur, "get_theta"):
            config_cur = config_cur.get_theta()
        return np.asarray(config_cur).reshape(shape)

    @staticmethod
    def _is_timedep(x):
        if hasattr(x, "__call__"):
            return True
        if len(np.array(x).shape) == 0:
            return False
        return np.array(x).shape[-1] > 1

    def _slice(self, x, time_step, N):
        if hasattr(x, "__call__"):
            return np.broadcast_to(x(time_step), (N, self.get_problem_count()))
        return x[:, :, time_step]

    def _f(self, time_step) -> np.ndarray:
        return self._slice(self.external_flux, time_step, self.circuit._Nf())

    def _Is(self, time_step) -> np.ndarray:
        return self._slice(self.current_sources, time_step, self.circuit._Nj())

    def _Vs(self, time_step) -> np.ndarray:
        return self._slice(self.voltage_sources, time_step, self.circuit._Nj())

    def _T(self, time_step) -> np.ndarray:
        return self._slice(self.temperature, time_step, self.circuit._Nj())

    def _cp(self, theta) -> np.ndarray:
        return self.current_phase_relation.eval(self.get_circuit()._Ic()[:, None], theta)

    def _dcp(self, theta) -> np.ndarray:
        return self.current_phase_relation.d_eval(self.get_circuit()._Ic()[:, None], theta)

    def _icp(self, theta) -> np.ndarray:
        return self.current_phase_relation.i_eval(self.get_circuit()._Ic()[:, None], theta)

    def _to_time_point_mask(self, time_points):
        if time_points is None:
            time_points = self.store_time_steps
        time_points = np.array(time_points)
        if time_points.dtype != bool:
            try:
                x = np.zeros(self._Nt(), dtype=bool)
                x[time_points] = True
                time_points = x
            except Exception:
                raise ValueError("Invalid store_time_steps; must be None, mask, slice or index array")
        return time_points

    def _get_stencil(self, width: int):
        if width == 3:
            return (1.0, -1.0, 0.0), (1.0, -2.0, 1.0)
        if width in (4, 5):
            raise NotImplementedError(
                "stencil_width 4 and 5 are not provided: in the reference they use a system matrix that is "
                "inconsistent with the stencil (flux quantisation violated at 3e-3, SURVEY.md quirk Q1)")
        raise ValueError(f"stencil width must be 3, 4 or 5 (equals {width})")


def _apply_derivative(x, index, stencil, dt):
    # reference: time_evolution.py:411-420 (3-point stencil)
    return (stencil[0] * x[:, :, index] + stencil[1] * x[:, :, index - 1]) / dt


def time_evolution(problem: TimeEvolutionProblem, core=None):
    """
    Decide which time points must be kept (voltage needs the preceding step as well), run the device
    core, finite-difference the voltage, trim helper steps. (reference: time_evolution.py:422-458)
    ``core`` (default: the GPU engine) has the signature of the reference's time_evolution_core.
    """
    Nt = problem._Nt()
    store = problem.store_time_steps
    zeros = np.zeros(Nt, dtype=bool)
    th_store_mask = store.copy() if (problem.store_theta or problem.store_voltage) else zeros.copy()
    I_store_mask = store.copy() if (problem.store_current or problem.store_voltage) else zeros.copy()
    V_th_store_mask = th_store_mask.copy()
    V_I_store_mask = I_store_mask.copy()
    t_ids = np.flatnonzero(store)
    Nj = problem.circuit._Nj()
    offset = problem.stencil_width - 1
    has_L = problem.circuit._has_inductance()
    if problem.store_voltage and len(t_ids) > 0:
        Vt_ids = (t_ids[:, None] - np.arange(offset)).ravel()
        Vt_ids = Vt_ids[(Vt_ids >= 0) & (Vt_ids < Nt)]
        V_th_store_mask[Vt_ids] = True
        if has_L:
            V_I_store_mask[Vt_ids] = True

    if core is None:
        # the initial-condition planes are only read by the voltage stencil below
        from .engine import device_time_evolution_core
        th_out, I_out = device_time_evolution_core(problem, V_th_store_mask, V_I_store_mask,
                                                   initial_planes=bool(problem.store_voltage))
    else:
        th_out, I_out = core(problem, V_th_store_mask, V_I_store_mask)

    V_out = None
    if problem.store_voltage:
        ts = np.flatnonzero(store[V_th_store_mask])
        V_out = _apply_derivative(th_out, index=ts + offset, stencil=problem.stencil[0], dt=problem._dt())
        if has_L:
            V_ind = _apply_derivative(I_out, index=ts + offset, stencil=problem.stencil[0], dt=problem._dt())
            V_out += (problem.circuit.get_inductance() @ V_ind.reshape((Nj, -1))).reshape(V_out.shape)
        th_out = np.delete(th_out, np.flatnonzero((V_th_store_mask & ~th_store_mask)[V_th_store_mask]) + offset, axis=2)
        I_out = np.delete(I_out, np.flatnonzero((V_I_store_mask & ~I_store_mask)[V_I_store_mask]) + offset, axis=2)
    th_out = th_out[:, :, offset:]
    I_out = I_out[:, :, offset:]
    return TimeEvolutionResult(problem, th_out if problem.store_theta else None,
                               I_out if problem.store_current else None,
                               V_out if problem.store_voltage else None)


class TimeEvolutionResult:
    """
    Data of simulated time evolution(s): theta, current, voltage of shape (Nj, W, Nt_s) (or None when
    not stored), plus quantities derived from them. (reference: time_evolution.py:597-1067)
    """

    def __init__(self, problem: TimeEvolutionProblem, theta, current, voltage):
        self.problem = problem
        Nj, W, Nt_s = problem.circuit._Nj(), self.get_problem_count(), problem._Nt_s()
        self.theta = theta
        self.voltage = voltage
        self.current = current
        for name, flag in (("theta", problem.store_theta), ("current", problem.store_current),
                           ("voltage", problem.store_voltage)):
            if flag:
                if getattr(self, name).shape != (Nj, W, Nt_s):
                    raise ValueError(f"{name} must have shape {(Nj, W, Nt_s)}; has shape {getattr(self, name).shape}")
            else:
                setattr(self, name, None)
        s = self.problem.store_time_steps.astype(int)
        self.time_point_indices = np.cumsum(s) - s
        self.animation = None

    def _th(self, time_point) -> np.ndarray:
        if self.theta is None:
            raise ThetaNotStored("Cannot query theta; quantity is not stored during time evolution simulation.")
        return self.theta[:, :, self._time_point_index(time_point)]

    def _V(self, time_point) -> np.ndarray:
        if self.voltage is None:
            raise VoltageNotStored("Cannot query voltage; quantity is not stored during time evolution simulation.")
        return self.voltage[:, :, self._time_point_index(time_point)]

    def _I(self, time_point) -> np.ndarray:
        if self.current is None:
            raise CurrentNotStored("Cannot query current; quantity is not stored during time evolution simulation.")
        return self.current[:, :, self._time_point_index(time_point)]

    def _time_point_index(self, time_points):
        if time_points is None:
            time_points = self.problem.store_time_steps
        if not np.all(self.problem.store_time_steps[time_points]):
            raise DataAtTimepointNotStored("Queried a timepoint that is not stored during time evolution simulation.")
        return self.time_point_indices[time_points]

    def get_problem_count(self):
        return self.problem.get_problem_count()

    def get_circuit(self) -> Circuit:
        return self.problem.get_circuit()

    def select_static_configuration(self, prob_nr, time_step):
        raise NotImplementedError("static configurations are outside the time-evolution hot path "
                                  "(reference: static_problem.py:815; SURVEY.md section 2 row C10)")

    def get_theta(self, select_time_points=None) -> np.ndarray:
        return self._select(select_time_points, self.get_circuit()._Nj(), self._th)

    def get_current(self, select_time_points=None) -> np.ndarray:
        return self._select(select_time_points, self.get_circuit()._Nj(), self._I)

    def get_voltage(self, select_time_points=None):
        return self._select(select_time_points, self.get_circuit()._Nj(), self._V)

    def get_phase(self, select_time_points=None) -> np.ndarray:
        c = self.get_circuit()
        M, Nj = c.get_cut_matrix(), c._Nj()
        func = lambda tp: c.Msq_solve(M @ self._th(tp).reshape(Nj, -1))
        try:
            return self._select(select_time_points, c._Nn(), func)
        except ThetaNotStored:
            raise ThetaNotStored("Cannot compute phi; requires theta to be stored in TimeEvolutionConfig")

    def get_vortex_configuration(self, select_time_points=None) -> np.ndarray:
        A = self.get_circuit().get_cycle_matrix()
        func = lambda tp: -A @ np.round(self._th(tp) / (2.0 * np.pi))
        try:
            return self._select(select_time_points, self.get_circuit()._Nf(), func).astype(int)
        except ThetaNotStored:
            raise ThetaNotStored("Cannot compute n; requires theta to be stored in TimeEvolutionConfig")

    def get_josephson_energy(self, select_time_points=None) -> np.ndarray:
        func = lambda tp: self.problem._icp(self._th(tp))
        try:
            return self._select(select_time_points, self.get_circuit()._Nj(), func)
        except ThetaNotStored:
            raise ThetaNotStored("Cannot compute Josephson energy EJ; requires theta to be stored in TimeEvolutionConfig")

    def get_supercurrent(self, select_time_points=None) -> np.ndarray:
        func = lambda tp: self.problem._cp(self._th(tp))
        try:
            return self._select(select_time_points, self.get_circuit()._Nj(), func)
        except ThetaNotStored:
            raise ThetaNotStored("Cannot compute supercurrent Isup; requires theta to be stored in TimeEvolutionConfig")

    def get_cycle_current(self, select_time_points=None) -> np.ndarray:
        A = self.get_circuit().get_cycle_matrix()
        func = lambda tp: self.get_circuit().Asq_solve(A @ (self._I(tp) - self.problem._Is(tp)))
        try:
            return self._select(select_time_points, self.get_circuit()._Nf(), func)
        except CurrentNotStored:
            raise CurrentNotStored("Cannot compute cycle-current J; requires current to be stored in TimeEvolutionConfig")

    def get_flux(self, select_time_points=None) -> np.ndarray:
        c = self.get_circuit()
        A = c.get_cycle_matrix()
        func = lambda tp: self.problem._f(tp) + A @ (c._L() @ self._I(tp)) / (2 * np.pi)
        try:
            return self._select(select_time_points, c._Nf(), func)
        except CurrentNotStored:
            raise CurrentNotStored("Cannot compute magnetic flux; requires current to be stored in TimeEvolutionConfig")

    def get_magnetic_energy(self, select_time_points=None) -> np.ndarray:
        c = self.get_circuit()
        func = lambda tp: 0.5 * c._L() @ (self._I(tp) ** 2)
        try:
            return self._select(select_time_points, c._Nj(), func, is_zero=not c._has_inductance())
        except CurrentNotStored:
            raise CurrentNotStored("Cannot compute magnetic energy EM; requires current to be stored in TimeEvolutionConfig")

    def get_potential(self, select_time_points=None):
        c = self.get_circuit()
        M, Nj = c.get_cut_matrix(), c._Nj()
        func = lambda tp: c.Msq_solve(M @ self._V(tp).reshape(Nj, -1))
        try:
            return self._select(select_time_points, c._Nn(), func)
        except VoltageNotStored:
            raise VoltageNotStored("Cannot compute electric potential U; requires voltage to be stored in TimeEvolutionConfig")

    def get_capacitive_energy(self, select_time_points=None):
        c = self.get_circuit()
        C = c._C()
        func = lambda tp: 0.5 * C[:, None] * self._V(tp) ** 2
        try:
            return self._select(select_time_points, c._Nj(), func, is_zero=not c._has_capacitance())
        except VoltageNotStored:
            raise VoltageNotStored("Cannot compute capacitive energy EC; requires voltage to be stored in TimeEvolutionConfig")

    def get_energy(self, select_time_points=None) -> np.ndarray:
        return self.get_josephson_energy(select_time_points) + self.get_magnetic_energy(select_time_points) + \
               self.get_capacitive_energy(select_time_points)

    def plot(self, *args, **kwargs):
        raise NotImplementedError("visualisation is outside the time-evolution hot path "
                                  "(reference: circuit_visualize.py; SURVEY.md section 2 row C13)")

    animate = plot

    def __str__(self):
        return "time evolution configuration: (" + \
               ("th" + str(self.theta.shape) + ", ") * (self.theta is not None) + \
               ("I" + str(self.current.shape) + ", ") * (self.current is not None) + \
               ("V" + str(self.voltage.shape)) * (self.voltage is not None) + ")" + \
               "\nproblem: " + self.problem.__str__() + \
               "\ncircuit: " + self.get_circuit().__str__()

    def _select(self, select_time_points, N, func, is_zero=False):
        select_time_points = np.flatnonzero(self.problem._to_time_point_mask(select_time_points))
        W = self.get_problem_count()
        out = np.zeros((N, W, len(select_time_points)), dtype=np.double)
        if is_zero:
            return out
        for i, tp in enumerate(select_time_points):
            out[:, :, i] = func(tp)
        return out


class AnnealedConfiguration:
    """
    One annealed problem: phases after the closing T = 0 runs and the vortex configuration the reference hands to
    its static solver (reference: time_evolution.py:1185-1188). The reference then polishes the phases with
    StaticProblem.compute() (static_problem.py) - a Newton solve outside the time-evolution path that this package
    does not provide; the object therefore carries the un-polished state, and ``AnnealingProblem.compute`` reports
    status 2 ("indeterminate") for it.
    """

    def __init__(self, circuit, theta, n, external_flux, current_sources):
        self.circuit, self.theta, self.n = circuit, theta, n
        self.external_flux, self.current_sources = external_flux, current_sources

    def get_circuit(self):
        return self.circuit

    def get_theta(self):
        return self.theta

    def get_n(self):
        return self.n

    def get_vortex_configuration(self):
        return self.n


class AnnealingProblem:
    """
    Anneals a circuit by gradually lowering the temperature; the temperature profile follows the measured vortex
    mobility (reference: time_evolution.py:1070-1191, same constructor, same schedule):

     - interval_count iterations of interval_steps time steps, the first at T = start_T; every iteration restarts
       from rest (theta(-2) = theta(-1));
     - after each iteration the vortex mobility sum |n(i+1) - n(i)| / (Nf dt (interval_count - 1)) per problem is
       compared with the target v (N - i)/N)^1.5 (or vortex_mobility[i]): above -> T /= T_factor, else T *= T_factor;
     - five closing runs at T = 0 with half the time step.

    The whole schedule runs on the GPU through ``engine.device_annealing``: state, factor and theta planes stay in
    HBM, per iteration one integer per problem comes back and one temperature per problem goes out. The reference
    re-enters compute() per iteration (refactorising, re-uploading and pulling interval_steps theta planes to the host).

    Extra keyword arguments: noise_seed, noise_replay ((interval_count, interval_steps, Nj, W) array or callable
    k -> (interval_steps, Nj, W)), devices.
    """

    def __init__(self, circuit: Circuit, time_step=0.5, interval_steps=10,
                 external_flux=0.0, current_sources=0, problem_count=1,
                 interval_count=1000, vortex_mobility=0.001,
                 start_T=1.0, T_factor=1.03, *, noise_seed=None, noise_replay=None, devices=None):
        self.circuit = circuit
        self.time_step = time_step
        self.interval_steps = interval_steps
        self.interval_count = interval_count
        self.vortex_mobility = vortex_mobility
        self.current_sources = current_sources
        self.external_flux = external_flux
        self.problem_count = problem_count
        self.T = start_T * np.ones((1, self.problem_count, 1))
        self.T_factor = T_factor
        self.noise_seed, self.noise_replay, self.devices = noise_seed, noise_replay, devices

    def get_vortex_mobility(self, n):
        """Vortex mobility of consecutive vortex configurations n (Nf, W, K) (reference: time_evolution.py:1128-1133)."""
        Nf = self.circuit.face_count()
        return np.sum(np.sum(np.abs(np.diff(n, axis=2)), axis=2), axis=0) / (Nf * self.time_step * (self.interval_count - 1))

    def _temperature_adjustment(self, vortex_mobility, iteration):
        # reference: time_evolution.py:1135-1140
        v = self.vortex_mobility
        upper = v[iteration] if (np.array(v)).size == self.interval_count else \
            v * ((self.interval_count - iteration) / self.interval_count) ** 1.5
        factor = (vortex_mobility > upper) * (1 / self.T_factor) + (vortex_mobility <= upper) * self.T_factor
        self.T *= factor[..., None]

    def _problem(self):
        # the problem the reference's loop re-runs (reference: time_evolution.py:1159-1162); constructing it validates
        # the inputs exactly as the reference does
        f = np.atleast_1d(self.external_flux)[:, None, None]
        return TimeEvolutionProblem(self.circuit, time_step_count=self.interval_steps, time_step=self.time_step,
                                    external_flux=f, current_sources=self.current_sources, temperature=self.T,
                                    store_current=False, store_voltage=False, stencil_width=3,
                                    noise_seed=self.noise_seed, noise_replay=self.noise_replay, devices=self.devices)

    def anneal(self):
        """
        The device part of compute(): the temperature schedule and the closing T = 0 runs.

        Returns
        -------
        theta : (Nj, problem_count) phases after the closing runs
        vortex_configuration : (Nf, problem_count) int array, n = -A round(theta / 2 pi)
        temperature_profiles : (interval_count, problem_count)
        """
        from .engine import device_annealing
        prob = self._problem()
        Nf, dt, N = self.circuit.face_count(), self.time_step, self.interval_count
        v, T_factor = self.vortex_mobility, self.T_factor

        def adjust(sums, i, T):
            # get_vortex_mobility + _temperature_adjustment on the exact integer sums of interval i
            mob = sums / (Nf * dt * (N - 1))
            upper = v[i] if (np.array(v)).size == N else v * ((N - i) / N) ** 1.5
            factor = (mob > upper) * (1 / T_factor) + (mob <= upper) * T_factor
            return T * factor

        out = device_annealing(prob, self.T[0, :, 0], adjust, N)
        self.T[0, :, 0] = out["T"]
        self.last_stats = out["stats"]
        return out["theta"], out["n"], out["profiles"]

    def compute(self):
        """
        Executes the annealing procedure (reference: time_evolution.py:1142-1191).

        Returns
        -------
        status : (problem_count,) int array; 2 (indeterminate) for every problem, because the static Newton solve
            with which the reference decides between 0 (converged) and 1 (diverged) is outside this package
        configurations : (problem_count,) list of AnnealedConfiguration (phases and vortex configuration of the
            annealed state, before the reference's static polish)
        temperature_profiles : (interval_count, problem_count) array
        """
        theta, n, profiles = self.anneal()
        f = np.atleast_1d(self.external_flux)
        configurations = [AnnealedConfiguration(self.circuit, theta[:, p].copy(), n[:, p].copy(), f, self.current_sources)
                          for p in range(self.problem_count)]
        status = np.full(self.problem_count, 2, dtype=int)
        return status, configurations, profiles
