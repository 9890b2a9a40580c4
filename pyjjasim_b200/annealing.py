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
        settings = dict(circuit=circuit, time_step=time_step, interval_steps=interval_steps, external_flux=external_flux,
                        current_sources=current_sources, problem_count=problem_count, interval_count=interval_count,
                        vortex_mobility=vortex_mobility, T_factor=T_factor, noise_seed=noise_seed,
                        noise_replay=noise_replay, devices=devices)
        for name, value in settings.items():
            setattr(self, name, value)
        self.T = np.full((1, problem_count, 1), float(start_T))        # one temperature per problem, (1, W, 1) like the reference's

    def mobility_targets(self):
        """(interval_count,) mobility target of every interval: vortex_mobility[i] when one value per interval is given,
        else vortex_mobility ((N - i) / N)^1.5 (reference: time_evolution.py:1136-1138)"""
        N, v = self.interval_count, self.vortex_mobility
        if np.array(v).size == N:
            return np.array([float(v[i]) for i in range(N)])
        return np.array([float(v * ((N - i) / N) ** 1.5) for i in range(N)])

    def mobility_norm(self):
        """what the integer mobility sums are divided by: Nf dt (interval_count - 1) (reference: time_evolution.py:1133)"""
        return self.circuit.face_count() * self.time_step * (self.interval_count - 1)

    def get_vortex_mobility(self, n):
        """Vortex mobility of consecutive vortex configurations n (Nf, W, K): the number of vortex moves per face and
        unit time (reference: time_evolution.py:1128-1133)."""
        moves = np.abs(n[:, :, 1:] - n[:, :, :-1]).sum(axis=2).sum(axis=0)
        return moves / self.mobility_norm()

    def _temperature_adjustment(self, vortex_mobility, iteration):
        """T /= T_factor for the problems whose mobility exceeds this interval's target, T *= T_factor for the others
        (reference: time_evolution.py:1135-1140)"""
        too_mobile = vortex_mobility > self.mobility_targets()[iteration]
        self.T *= np.where(too_mobile, 1 / self.T_factor, self.T_factor)[..., None]

    def _problem(self):
        # the problem the reference's loop re-runs (reference: time_evolution.py:1159-1162); constructing it validates
        # the inputs exactly as the reference does
        from .time_evolution import TimeEvolutionProblem
        return TimeEvolutionProblem(self.circuit, time_step=self.time_step, time_step_count=self.interval_steps,
                                    external_flux=np.atleast_1d(self.external_flux)[:, None, None],
                                    current_sources=self.current_sources, temperature=self.T, stencil_width=3,
                                    store_current=False, store_voltage=False,
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
        N, T_factor = self.interval_count, self.T_factor
        targets, norm = self.mobility_targets(), self.mobility_norm()

        def adjust(sums, i, T):
            # the rule of _temperature_adjustment on the exact integer sums of interval i (host path: replayed noise)
            return T * np.where(sums / norm > targets[i], 1 / T_factor, T_factor)

        # the same rule as data, for the device-side schedule (jj_anneal)
        rule = dict(upper=targets, T_factor=float(T_factor), norm=float(norm)) if N > 1 else None
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
