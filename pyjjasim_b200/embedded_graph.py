"""
Planar embedded graphs: nodes with coordinates, edges, and the faces they enclose.

Host-side setup code (runs once per circuit). It supplies the two incidence
matrices the time-evolution hot path consumes:

* the face cycle matrix ``A`` (Nf, Nj), entries +-1  (reference: embedded_graph.py:525-539)
* the cut matrix ``M`` (Nn, Nj), entries +-1         (reference: embedded_graph.py:541-554)

Index conventions are those of the reference so that inputs given per junction or
per face mean the same thing (checked against the reference in tests/golden):

* junction ``j`` is edge ``j`` in the order the user passed them, directed node1 -> node2;
* ``A[f, j] = +1`` if walking face ``f`` counter-clockwise passes junction ``j`` in its own direction;
* faces are ordered by (number of edges, smallest "sorted" edge id they contain), where the
  sorted edge order is lexicographic in (min node, max node)  (reference: embedded_graph.py:861-867
  for the edge order, :1009-1036 for the order in which cycles are emitted).

The algorithm here is not the reference's breadth-wise cycle walk: faces are found by
building the half-edge successor permutation and labelling its orbits with pointer doubling
(O(log(longest cycle)) vectorised passes).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse

__all__ = ["EmbeddedGraph", "EmbeddedSquareGraph", "EmbeddedHoneycombGraph",
           "EmbeddedTriangularGraph", "NotPlanarEmbeddingError", "NotSingleComponentError",
           "SelfLoopError", "NonSimpleError"]


class NotSingleComponentError(Exception):
    pass


class NotPlanarEmbeddingError(Exception):
    pass


class SelfLoopError(Exception):
    pass


class NonSimpleError(Exception):
    pass


class EmbeddedGraph:
    """
    Embedded 2D graph.

    Parameters
    ----------
    x, y : (N,) arrays
        node coordinates.
    node1, node2 : (E,) int arrays in range(N)
        end points of each edge; edge e is directed node1[e] -> node2[e].
    """

    def __init__(self, x, y, node1, node2, require_single_component=False,
                 require_planar_embedding=False):
        self.x = np.array(x, dtype=np.double).ravel()
        self.y = np.array(y, dtype=np.double).ravel()
        if self.x.size != self.y.size:
            raise ValueError("x and y must be same size")
        self.node1 = np.array(node1, dtype=np.int64).ravel()
        self.node2 = np.array(node2, dtype=np.int64).ravel()
        if self.node1.size != self.node2.size:
            raise ValueError("node1 and node2 must be same size")
        N, E = self.x.size, self.node1.size
        if E and (min(self.node1.min(), self.node2.min()) < 0 or
                  max(self.node1.max(), self.node2.max()) >= N):
            raise ValueError("edges refer to non-existing nodes")
        if np.any(self.node1 == self.node2):
            raise SelfLoopError("graph contains self-loops")
        lo, hi = np.minimum(self.node1, self.node2), np.maximum(self.node1, self.node2)
        if np.unique(lo * N + hi).size != E:
            raise NonSimpleError("graph is not simple (duplicate edges)")
        # rank of each edge in the (min node, max node) lexicographic order
        order = np.lexsort((hi, lo))
        self._sorted_rank = np.empty(E, dtype=np.int64)
        self._sorted_rank[order] = np.arange(E)
        self._faces = None
        self._n_components = None
        if require_single_component:
            self._assert_single_component()
        if require_planar_embedding:
            self._assert_planar_embedding()

    # ------------------------------------------------------------------ basic queries
    def node_count(self):
        return self.x.size

    def edge_count(self):
        return self.node1.size

    def face_count(self):
        return len(self._get_faces()["length"])

    def coo(self):
        return self.x, self.y

    def get_edges(self):
        return self.node1, self.node2

    def get_num_components(self):
        if self._n_components is None:
            N, E = self.node_count(), self.edge_count()
            adj = scipy.sparse.coo_matrix((np.ones(E), (self.node1, self.node2)), shape=(N, N))
            self._n_components = scipy.sparse.csgraph.connected_components(adj, directed=False)[0]
        return self._n_components

    def is_planar_embedding(self):
        # Euler's formula counting only positively oriented face cycles, as the reference does
        # (reference: embedded_graph.py:404-409): crossing edges produce cycles of non-positive area.
        return self.node_count() + self.face_count() == self.edge_count() + 1

    def _assert_single_component(self):
        if self.get_num_components() != 1:
            raise NotSingleComponentError("graph is not single-component")

    def _assert_planar_embedding(self):
        if not self.is_planar_embedding():
            raise NotPlanarEmbeddingError("graph is not a planar embedding")

    # ------------------------------------------------------------------ faces
    def _successor(self):
        """Half-edge h in [0, 2E): h < E is node1->node2 of edge h, h >= E its reverse.
        succ[h] = the half-edge leaving head(h) that makes the left-most turn."""
        E = self.edge_count()
        tail = np.concatenate((self.node1, self.node2))
        head = np.concatenate((self.node2, self.node1))
        ang = np.arctan2(self.y[head] - self.y[tail], self.x[head] - self.x[tail])
        # sort outgoing half-edges per tail node by angle (counter-clockwise)
        order = np.lexsort((ang, tail))
        pos = np.empty(2 * E, dtype=np.int64)
        pos[order] = np.arange(2 * E)
        start = np.searchsorted(tail[order], np.arange(self.node_count()))
        deg = np.bincount(tail, minlength=self.node_count())
        # the reverse of h leaves head(h); its clockwise neighbour is the left-most turn
        rev = np.concatenate((np.arange(E, 2 * E), np.arange(E)))
        p = pos[rev]
        v = head
        prev_pos = start[v] + (p - start[v] - 1) % deg[v]
        return order[prev_pos]

    def _get_faces(self):
        if self._faces is not None:
            return self._faces
        E = self.edge_count()
        H = 2 * E
        succ = self._successor()
        # orbit labels = smallest half-edge id on the orbit, by pointer doubling
        label = np.arange(H)
        jump = succ.copy()
        # after k rounds label[h] = min over the 2^k half-edges following h; a labelling that is
        # constant along succ is constant on orbits and then equals the orbit minimum.
        while H and not np.array_equal(label[succ], label):
            label = np.minimum(label, label[jump])
            jump = jump[jump]
        uniq, inv, length = np.unique(label, return_inverse=True, return_counts=True)
        n_orb = uniq.size
        tail = np.concatenate((self.node1, self.node2))
        head = np.concatenate((self.node2, self.node1))
        # shoelace area of every orbit
        cross = self.x[tail] * self.y[head] - self.x[head] * self.y[tail]
        area = 0.5 * np.bincount(inv, weights=cross, minlength=n_orb)
        # ordering key: smallest sorted-edge rank among the orbit's forward (low->high node) half-edges
        fwd = tail < head
        edge_of = np.arange(H) % E
        key = np.full(n_orb, np.iinfo(np.int64).max)
        np.minimum.at(key, inv[fwd], self._sorted_rank[edge_of[fwd]])
        interior = ~((area < 0) | np.isclose(area, 0.0))
        face_orb = np.flatnonzero(interior)
        face_orb = face_orb[np.lexsort((key[face_orb], length[face_orb]))]
        face_id = np.full(n_orb, -1, dtype=np.int64)
        face_id[face_orb] = np.arange(face_orb.size)
        he_face = face_id[inv]            # face of each half-edge (-1: boundary orbit)
        # centroid of each face (mean of its corner nodes) - used for geometric orderings
        cnt = np.maximum(length, 1)
        cx = np.bincount(inv, weights=self.x[tail], minlength=n_orb) / cnt
        cy = np.bincount(inv, weights=self.y[tail], minlength=n_orb) / cnt
        self._faces = dict(he_face=he_face, length=length[face_orb], area=area[face_orb],
                           cx=cx[face_orb], cy=cy[face_orb], orbit_count=n_orb, succ=succ)
        return self._faces

    def get_areas(self):
        return self._get_faces()["area"]

    def get_face_centroids(self):
        f = self._get_faces()
        return f["cx"], f["cy"]

    def face_cycle_matrix(self):
        """(Nf, Nj) int64 CSC matrix, see module docstring. (reference: embedded_graph.py:525-539)"""
        f = self._get_faces()
        E, F = self.edge_count(), len(f["length"])
        he_face = f["he_face"]
        h = np.flatnonzero(he_face >= 0)
        rows, cols = he_face[h], h % E
        data = np.where(h < E, 1, -1).astype(np.int64)
        m = scipy.sparse.coo_matrix((data, (rows, cols)), shape=(F, E)).tocsc()
        m.sum_duplicates()
        m.sort_indices()
        return m

    def cut_space_matrix(self):
        """(Nn, Nj) CSC matrix: -1 at node1, +1 at node2. (reference: embedded_graph.py:541-554)"""
        E, N = self.edge_count(), self.node_count()
        row = np.concatenate((self.node1, self.node2))
        col = np.concatenate((np.arange(E), np.arange(E)))
        data = np.concatenate((-np.ones(E), np.ones(E)))
        return scipy.sparse.coo_matrix((data, (row, col)), shape=(N, E)).tocsc()


class EmbeddedSquareGraph(EmbeddedGraph):
    """Square lattice with count_x by count_y nodes; horizontal edges first (row-major), then vertical.
    (reference: embedded_graph.py:1093-1111)"""

    def __init__(self, count_x, count_y, x_scale=1.0, y_scale=1.0):
        iy, ix = np.divmod(np.arange(count_x * count_y), count_x)
        ids = np.arange(count_x * count_y).reshape(count_y, count_x)
        n1 = np.concatenate((ids[:, :-1].ravel(), ids[:-1, :].ravel()))
        n2 = np.concatenate((ids[:, 1:].ravel(), ids[1:, :].ravel()))
        super().__init__(ix * x_scale, iy * y_scale, n1, n2)


class EmbeddedHoneycombGraph(EmbeddedGraph):
    """Honeycomb lattice of count_x by count_y four-node unit cells. (reference: embedded_graph.py:1114-1139)"""

    def __init__(self, count_x, count_y, x_scale=1.0, y_scale=1.0):
        s = count_x * count_y
        iy, ix = np.divmod(np.arange(s), count_x)
        bx, by = 3.0 * ix, np.sqrt(3.0) * iy
        h = np.sqrt(0.75)
        xs = np.concatenate((bx, bx + 0.5, bx + 1.5, bx + 2.0))
        ys = np.concatenate((by, by + h, by + h, by))
        ids = np.arange(s).reshape(count_y, count_x)
        a, b, c, d = ids, ids + s, ids + 2 * s, ids + 3 * s   # the four nodes of each cell
        pairs = [(a, b), (b[:-1, :], a[1:, :]), (b, c), (c[:-1, :], d[1:, :]), (c, d), (d[:, :-1], a[:, 1:])]
        n1 = np.concatenate([p[0].ravel() for p in pairs])
        n2 = np.concatenate([p[1].ravel() for p in pairs])
        super().__init__(xs * x_scale, ys * y_scale, n1, n2)


class EmbeddedTriangularGraph(EmbeddedGraph):
    """Triangular lattice of count_x by count_y two-node unit cells. (reference: embedded_graph.py:1142-1166)"""

    def __init__(self, count_x, count_y, x_scale=1.0, y_scale=1.0):
        s = count_x * count_y
        iy, ix = np.divmod(np.arange(s), count_x)
        bx, by = 1.0 * ix, np.sqrt(3.0) * iy
        xs = np.concatenate((bx, bx + 0.5))
        ys = np.concatenate((by, by + np.sqrt(0.75)))
        ids = np.arange(s).reshape(count_y, count_x)
        a, b = ids, ids + s
        pairs = [(a, b), (a[:, :-1], a[:, 1:]), (b[:-1, :], a[1:, :]), (b[:-1, :-1], a[1:, 1:]),
                 (b[:, :-1], b[:, 1:]), (a[:, 1:], b[:, :-1])]
        n1 = np.concatenate([p[0].ravel() for p in pairs])
        n2 = np.concatenate([p[1].ravel() for p in pairs])
        super().__init__(xs * x_scale, ys * y_scale, n1, n2)
