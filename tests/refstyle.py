"""TEST INFRASTRUCTURE - stand-ins shaped like the REFERENCE's own problem and circuit objects.

INTEGRATION.md binds `device_time_evolution_core` to the reference's `TimeEvolutionProblem`. Such an object differs
from this package's: it has no `_raw_sources`; every non-callable input is held as an (N, W, Nt) read-only broadcast
view (reference: time_evolution.py:117-131) with the time-dependence flags frozen at construction (:334-340); and the
circuit offers the reference's getters only. The reference tree is absent on the GPU box, so the GPU test drives this
stand-in; tests/test_host.py checks, where /root/reference exists, that the stand-in and the real class hold the same
attributes in the same form."""
import numpy as np


class RefStyleCircuit:
    """The getters of the reference's Circuit that the hot path may touch, nothing else (no `_raw` anything)."""

    _GETTERS = ("_Nj", "_Nf", "_Nn", "_Ic", "_R", "_C", "_L", "_has_inductance", "get_cycle_matrix", "get_cut_matrix",
                "get_node_coordinates", "get_junction_nodes", "get_face_centroids")

    def __init__(self, circuit):
        for name in self._GETTERS:
            if hasattr(circuit, name):
                setattr(self, name, getattr(circuit, name))


def _timedep(x):
    if callable(x):
        return True
    shape = np.array(x).shape
    return len(shape) > 0 and shape[-1] > 1


def _count(x):
    shape = np.array(x(0) if callable(x) else x).shape
    return shape[1] if len(shape) > 1 else 1


class RefStyleProblem:
    def __init__(self, circuit, time_step=0.05, time_step_count=1000, current_phase_relation=None, external_flux=0.0,
                 current_sources=0.0, voltage_sources=0.0, temperature=0.0, config_at_minus_1=None,
                 config_at_minus_2=None, stencil_width=3):
        self.circuit, self.time_step, self.time_step_count = circuit, time_step, time_step_count
        self.current_phase_relation = current_phase_relation
        self.stencil_width = stencil_width
        inputs = dict(external_flux=external_flux, current_sources=current_sources, voltage_sources=voltage_sources,
                      temperature=temperature)
        self.problem_count = max(_count(v) for v in inputs.values())
        Nj, Nf, W, Nt = circuit._Nj(), circuit._Nf(), self.problem_count, time_step_count
        for name, flag, N in (("external_flux", "_f_is_timedep", Nf), ("current_sources", "_Is_is_timedep", Nj),
                              ("voltage_sources", "_Vs_is_timedep", Nj), ("temperature", "_T_is_timedep", Nj)):
            v = inputs[name]
            setattr(self, flag, _timedep(v))
            setattr(self, name, v if callable(v) else np.broadcast_to(np.array(v), (N, W, Nt)))
        m1 = np.zeros((Nj, W)) if config_at_minus_1 is None else np.asarray(config_at_minus_1).reshape(Nj, W)
        self.config_at_minus_1 = m1
        self.config_at_minus_2 = m1.copy() if config_at_minus_2 is None else np.asarray(config_at_minus_2).reshape(Nj, W)

    def get_circuit(self):
        return self.circuit

    def get_problem_count(self):
        return self.problem_count

    def _Nt(self):
        return self.time_step_count

    def _dt(self):
        return self.time_step

    def _cp(self, theta):
        Ic = self.circuit._Ic()
        return self.current_phase_relation.eval(Ic[:, None] if theta.ndim > 1 else Ic, theta)
