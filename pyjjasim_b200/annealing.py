"""
``AnnealingProblem``: the annealing caller of the stepping loop (reference: time_evolution.py:1070-1191, same
constructor, same schedule), run as ONE device-resident schedule instead of one compute() per interval.
"""
import numpy as np

from .josephson_circuit import Circuit

__all__ = ["AnnealingProblem", "AnnealedConfiguration"]


from .static_polish import AnnealedConfiguration, london_approximation, newton_stationary_states, _CircuitOps   # noqa: E402


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
        from .time_evolution import TimeEvolutionProblem
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

        # the same rule as data, for the device-side schedule
        upper_all = np.array([float(v[i]) if (np.array(v)).size == N else float(v * ((N - i) / N) ** 1.5) for i in range(N)])
        rule = dict(upper=upper_all, T_factor=float(T_factor), norm=float(Nf * dt * (N - 1))) if N > 1 else None
        out = device_annealing(prob, self.T[0, :, 0], adjust, N, rule=rule)
        self.T[0, :, 0] = out["T"]
        self.last_stats = out["stats"]
        return out["theta"], out["n"], out["profiles"]

    def compute(self, polish=True):
        """
        Executes the annealing procedure (reference: time_evolution.py:1142-1191): the temperature schedule and the
        closing T = 0 runs on the GPU (``anneal``), then - like the reference - one stationary-state solve per problem
        for the annealed vortex configuration, started from the London approximation (host, static_polish.py).

        Returns
        -------
        status : (problem_count,) int array: 0 converged onto the annealed vortex configuration, 1 diverged or
            converged onto another one, 2 iteration limit (polish=False: 2 for every problem, nothing is solved)
        configurations : (problem_count,) list of AnnealedConfiguration (stationary phases, vortex configuration,
            getters of the reference's StaticConfiguration; ``annealed_theta`` holds the phases the anneal ended with)
        temperature_profiles : (interval_count, problem_count) array
        """
        theta, n, profiles = self.anneal()
        prob = self._problem()
        W, cpr = self.problem_count, prob.current_phase_relation
        f = np.array(np.broadcast_to(prob._f(0), (self.circuit._Nf(), W)), dtype=np.double)
        Is = np.array(np.broadcast_to(prob._Is(0), (self.circuit._Nj(), W)), dtype=np.double)
        if polish:
            ops = _CircuitOps(self.circuit)
            theta0 = london_approximation(self.circuit, f, n, Is, ops)
            theta_s, status, info = newton_stationary_states(self.circuit, theta0, Is, f, n, cpr, ops=ops)
            self.last_polish = info
        else:
            theta_s, status, info = theta, np.full(W, 2, dtype=int), dict(error=np.full(W, np.nan))
        configurations = [AnnealedConfiguration(self.circuit, theta_s[:, p].copy(), n[:, p].copy(), f[:, p], Is[:, p], cpr,
                                                annealed_theta=theta[:, p].copy(), error=float(info["error"][p]))
                          for p in range(W)]
        return status, configurations, profiles
