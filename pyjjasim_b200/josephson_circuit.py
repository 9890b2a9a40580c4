"""
Josephson circuits: the *input object* of the time-evolution hot path.

Only the parts of the reference's ``Circuit`` that the path reads are provided
(reference: josephson_circuit.py:22-1020):

* the cycle matrix ``A`` and cut matrix ``M``,
* per-junction ``Ic``, ``R``, ``C`` vectors and the (Nj, Nj) inductance matrix ``L``,
* the lattice constructors ``SquareArray``, ``HoneycombArray``, ``TriangularArray``, ``SQUID``
  with the reference's index conventions and ``current_base(angle)``.

Any object exposing the same getters (for example the reference's own ``Circuit``) can be
passed to ``TimeEvolutionProblem`` instead: the path is duck-typed on
``get_cycle_matrix, _R, _C, _Ic, _L, _Nj, _Nf, _has_inductance``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

from .embedded_graph import (EmbeddedGraph, EmbeddedSquareGraph, EmbeddedHoneycombGraph,
                             EmbeddedTriangularGraph)

__all__ = ["Circuit", "Lattice", "SquareArray", "HoneycombArray", "TriangularArray", "SQUID",
           "node_to_junction_current"]


class Circuit:
    """
    Josephson junction array on a planar embedded graph; every edge holds one junction.
    (reference: josephson_circuit.py:22-91)

    Parameters
    ----------
    graph : EmbeddedGraph
    critical_current, resistance, capacitance : scalar or (Nj,) array
    inductance : scalar, (Nj,) array (self inductance) or symmetric (Nj, Nj) matrix
    """

    def __init__(self, graph: EmbeddedGraph, critical_current=1.0, resistance=1.0,
                 capacitance=0.0, inductance=0.0, negative_Ic_allowed=False):
        self.graph = graph
        graph._assert_planar_embedding()
        graph._assert_single_component()
        self.negative_Ic_allowed = negative_Ic_allowed
        self.cut_matrix = graph.cut_space_matrix()
        self.cycle_matrix = graph.face_cycle_matrix()
        self._Msq_factorized = None
        self._Asq_factorized = None
        self.set_resistance(resistance)
        self.set_critical_current(critical_current)
        self.set_capacitance(capacitance)
        self.set_inductance(inductance)

    # --- geometry -------------------------------------------------------------------
    def get_node_coordinates(self):
        return self.graph.coo()

    def get_junction_nodes(self):
        return self.graph.get_edges()

    def node_count(self):
        return self.graph.node_count()

    def junction_count(self):
        return self.graph.edge_count()

    def face_count(self):
        return self.graph.face_count()

    def get_face_centroids(self):
        return self.graph.get_face_centroids()

    def get_face_areas(self):
        return self.graph.get_areas()

    def _Nn(self):
        return self.graph.node_count()

    def _Nj(self):
        return self.graph.edge_count()

    def _Nf(self):
        return self.graph.face_count()

    # --- component values (reference: josephson_circuit.py:400-461, 582-604, 701-711) ---
    @staticmethod
    def _junction_quantity(x, N, name):
        try:
            return np.broadcast_to(x, (N,)).copy().astype(np.double)
        except ValueError:
            raise ValueError(name + " must be scalar or array of length equal to junction count")

    def get_critical_current(self):
        return self.critical_current

    def set_critical_current(self, Ic):
        Ic = self._junction_quantity(Ic, self._Nj(), "Ic")
        if not self.negative_Ic_allowed and np.any(Ic < 0):
            raise ValueError("Defined negative critical current. If intentional;"
                             "set negative_Ic_allowed field of circuit to True.")
        self.critical_current = Ic
        return self

    def get_resistance(self):
        return self.resistance

    def set_resistance(self, R):
        R = self._junction_quantity(R, self._Nj(), "R")
        if np.any(R <= 0.0):
            raise ValueError("All junctions must have a positive resistor")
        self.resistance = R
        return self

    def get_capacitance(self):
        return self.capacitance

    def set_capacitance(self, C):
        C = self._junction_quantity(C, self._Nj(), "C")
        if np.any(C < 0.0):
            raise ValueError("Capacitance cannot be negative.")
        self.capacitance = C
        return self

    def get_inductance(self):
        return self.inductance

    def set_inductance(self, inductance):
        N = self._Nj()
        L = inductance if hasattr(inductance, "ndim") else np.array(inductance)
        if L.ndim <= 1:
            v = self._junction_quantity(L, N, "L")
            if not np.all(v > -10 * np.finfo(float).eps):
                raise ValueError("Inductance matrix not positive definite")
            self.inductance = scipy.sparse.diags(v, 0).tocsc()
            self._has_inductance_v = bool(np.any(v != 0.0))
            return self
        if L.shape != (N, N):
            raise ValueError("L must be scalar, (Nj,) array or (Nj, Nj) matrix")
        Ls = scipy.sparse.csc_matrix(L)
        if (Ls - Ls.T).nnz != 0:
            raise ValueError("inductance matrix must be symmetric")
        if Ls.nnz:
            # positive definiteness through the smallest eigenvalue bound of a Cholesky-like LU
            lu = scipy.sparse.linalg.splu(Ls, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                                          options=dict(SymmetricMode=True))
            if not (np.all(lu.perm_r == lu.perm_c) and np.all(lu.U.diagonal() > 0)):
                raise ValueError("Inductance matrix not positive definite")
        self.inductance = Ls
        # Reference quirk, reproduced for drop-in parity: for a matrix-valued inductance the reference
        # stores "matrix is zero" in its has-inductance flag (josephson_circuit.py:733-755 returns
        # (L, is_positive_definite, is_zero) into (inductance, _, _has_inductance_v) at :602-603), so a
        # non-zero mutual-inductance matrix reports _has_inductance() == False. The stepping loop always
        # uses L (time_evolution.py:504-505); only the post-hoc inductive voltage term (:448-450) is skipped.
        self._has_inductance_v = (Ls.nnz == 0)
        return self

    def _Ic(self):
        return self.critical_current

    def _R(self):
        return self.resistance

    def _C(self):
        return self.capacitance

    def _L(self):
        return self.inductance

    def _has_capacitance(self):
        return bool(np.any(self.capacitance > 0))

    def _has_inductance(self):
        return self._has_inductance_v

    # --- incidence matrices (reference: josephson_circuit.py:606-625) -----------------
    def get_cut_matrix(self):
        return self.cut_matrix

    def get_cycle_matrix(self):
        return self.cycle_matrix

    def _Mr(self):
        return self.cut_matrix[:-1, :]

    def Msq_solve(self, b):
        """Solve M M^T x = b with the last node grounded. (reference: josephson_circuit.py:93-109)"""
        if self._Msq_factorized is None:
            Mr = self._Mr()
            self._Msq_factorized = scipy.sparse.linalg.factorized((Mr @ Mr.T).tocsc())
        return np.append(self._Msq_factorized(b[:-1, ...]), np.zeros(b[-1:, ...].shape), axis=0)

    def Asq_solve(self, b):
        """Solve A A^T x = b. (reference: josephson_circuit.py:111-128)"""
        if self._Asq_factorized is None:
            A = self.cycle_matrix
            self._Asq_factorized = scipy.sparse.linalg.factorized((A @ A.T).tocsc().astype(np.double))
        return self._Asq_factorized(b)

    def __str__(self):
        return "(x, y): \n" + str(np.stack(self.get_node_coordinates())) + \
               "\n(node1_id, node2_id): \n" + str(np.stack(self.get_junction_nodes()))


def node_to_junction_current(circuit: Circuit, node_current):
    """Junction currents whose divergence equals the given node currents.
    (reference: static_problem.py:1174-1192)"""
    return - circuit.get_cut_matrix().T @ circuit.Msq_solve(node_current)


class Lattice:
    """Shared behaviour of the regular arrays. (reference: josephson_circuit.py:893-947)"""

    def __init__(self, count_x, count_y, x_scale=1.0, y_scale=1.0):
        self.count_x, self.count_y = count_x, count_y
        self.x_scale, self.y_scale = x_scale, y_scale

    def get_count_x(self):
        return self.count_x

    def get_count_y(self):
        return self.count_y

    def get_x_scale(self):
        return self.x_scale

    def get_y_scale(self):
        return self.y_scale

    def current_base(self, angle, type="junction"):
        """Current sources producing a uniform current at the given angle (radians)."""
        if type == "junction":
            return node_to_junction_current(self, self.current_base(angle, type="node"))
        if type == "node":
            return self._I_node_h() * np.cos(angle) + self._I_node_v() * np.sin(angle)
        raise ValueError("type must be 'junction' or 'node'")


class SquareArray(Circuit, Lattice):
    """(reference: josephson_circuit.py:952-966)"""

    def __init__(self, count_x, count_y, x_scale=1.0, y_scale=1.0):
        Lattice.__init__(self, count_x, count_y, x_scale, y_scale)
        Circuit.__init__(self, EmbeddedSquareGraph(count_x, count_y, x_scale, y_scale))

    def _I_node_h(self):
        x, _ = self.get_node_coordinates()
        return (x == 0).astype(int) - np.isclose(x, (self.count_x - 1) * self.x_scale).astype(int)

    def _I_node_v(self):
        _, y = self.get_node_coordinates()
        return (y == 0).astype(int) - np.isclose(y, (self.count_y - 1) * self.y_scale).astype(int)


class HoneycombArray(Circuit, Lattice):
    """(reference: josephson_circuit.py:969-983)"""

    def __init__(self, count_x, count_y, x_scale=1.0, y_scale=1.0):
        Lattice.__init__(self, count_x, count_y, x_scale, y_scale)
        Circuit.__init__(self, EmbeddedHoneycombGraph(count_x, count_y, x_scale, y_scale))

    def _I_node_h(self):
        x, _ = self.get_node_coordinates()
        n = self.count_y / (self.count_y - 0.5)
        return ((x == 0).astype(int) - np.isclose(x, (3 * self.count_x - 1) * self.x_scale).astype(int)) * n

    def _I_node_v(self):
        _, y = self.get_node_coordinates()
        return ((y < 0.1 * np.sqrt(3) * self.y_scale).astype(int) -
                (y > (self.count_y - 0.6) * np.sqrt(3) * self.y_scale).astype(int)) * 1.5


class TriangularArray(Circuit, Lattice):
    """(reference: josephson_circuit.py:985-999)"""

    def __init__(self, count_x, count_y, x_scale=1.0, y_scale=1.0):
        Lattice.__init__(self, count_x, count_y, x_scale, y_scale)
        Circuit.__init__(self, EmbeddedTriangularGraph(count_x, count_y, x_scale, y_scale))

    def _I_node_h(self):
        x, _ = self.get_node_coordinates()
        return ((x < 0.1 * self.x_scale).astype(int) -
                (x > (self.count_x - 0.6) * self.x_scale).astype(int)) * np.sqrt(3)

    def _I_node_v(self):
        _, y = self.get_node_coordinates()
        return (y == 0).astype(int) - np.isclose(y, (self.count_y - 0.5) * np.sqrt(3) * self.y_scale).astype(int)


class SQUID(Circuit):
    """A square whose vertical junctions have Ic=1000 and horizontal ones Ic=1.
    (reference: josephson_circuit.py:1001-1020)"""

    def __init__(self):
        graph = EmbeddedGraph([0, 1, 1, 0], [0, 0, 1, 1], [0, 1, 2, 0], [1, 2, 3, 3])
        super().__init__(graph, critical_current=[1, 1000, 1, 1000])

    def horizontal_junctions(self):
        return np.array([1, 0, -1, 0])

    def vertical_junctions(self):
        return np.array([0, 1, 0, 1])
