"""
Host-side factorisation of the cycle-space system and its compilation into *solve programs* for
the device engines.

Reference behaviour being replaced: ``scipy.sparse.linalg.factorized(A_mat)`` once
(reference: time_evolution.py:504-506, SuperLU with COLAMD) and two SuperLU triangular sweeps per
time step (reference: time_evolution.py:560-569).

Here (all on the host, once per problem):
  1. S = A (L + diag(1/(Cv+Rv))) A^T is permuted with a geometric nested dissection (ordering.py);
  2. the permuted matrix is factorised S = Lc Lc^T (SuperLU in symmetric mode without pivoting,
     rescaled to a Cholesky factor);
  3. the factor is cut along the dissection blocks. For a block B with diagonal block D = Lc[B, B]
     the device never does a sequential triangular solve; it applies the explicit inverse D^-1:
        forward    t_B = b_B - Lc[B, <B] z_{<B}     (phase a: sparse rows, independent)
                   z_B = D^-1 t_B                    (phase b: dense lower-triangular rows, independent)
        backward   t_B = z_B - Lc[>B, B]^T J_{>B}   (phase a)
                   J_B = D^-T t_B                    (phase b)
     so one sweep has 2*height+1 dependent phases instead of one per elimination level.
  4. every phase is emitted as a list of *tiles*: a tile is a small dense matrix V (nrows x ncols,
     nrows * lanes_per_row == 32) with a column index list; out[row0 + i] = (self ? src[row0 + i] : 0)
     + sum_c V[i, c] * src[cols[c]]. Values are stored in the order a warp consumes them
     ([step][lane], lane = i * lanes_per_row + c % lanes_per_row), so device loads are coalesced.

``streaming_program`` packs these tiles with permuted face indices as rows (streaming engine, any size, any
input form); the fast path (subdomain engine) has its own plan builder in subdomain.py, which reads the
factor ``Lc`` and the dissection tree kept in ``Factor``.
"""
import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

from .ordering import nested_dissection

__all__ = ["Factor", "factorize", "streaming_program", "system_matrix", "apply_program_host",
           "build_solve_program"]

TILE_SELF = 1        # add src[row] to the dot product (phase a)
TILE_STAGED = 2      # block spans several tiles: results must not overwrite src before the level is complete



def system_matrix(A, Lmat, Rv, Cv):
    """A (L + diag(1/(Cv+Rv))) A^T as CSC float64. (reference: time_evolution.py:476,504-505)"""
    A = scipy.sparse.csc_matrix(A).astype(np.double)
    mid = scipy.sparse.csc_matrix(Lmat) + scipy.sparse.diags(1.0 / (Cv + Rv), 0)
    return scipy.sparse.csc_matrix(A @ mid @ A.T)


class Factor:
    """Block Cholesky factor of the permuted system: S[perm][:, perm] = Lc Lc^T."""

    def __init__(self):
        self.n = 0
        self.perm = None          # new-to-old
        self.bptr = None          # block b owns permuted rows bptr[b]:bptr[b+1]
        self.height = None        # leaves 0
        self.depth = None         # cut-tree depth of the block
        self.dom = None           # cut-tree path of the block
        self.dinv = None          # list: inverse of the diagonal block (dense lower triangular)
        self._Loff = None         # CSR: Lc without the diagonal blocks  (row i: couplings to earlier blocks)
        self._LoffT = None        # CSR: its transpose                   (row j: couplings to later blocks)
        self.nnz_L = 0
        self.blk_part = None      # part (subtree) of each block when the ordering was made with n_parts
        self.Lc = None            # CSR: the full Cholesky factor (subdomain engine: Schur complement of the top separators)
        self.Sp = None            # CSR: the permuted system matrix

    @property
    def Loff(self):
        if self._Loff is None:
            blk_of = np.repeat(np.arange(self.nb), np.diff(self.bptr))
            coo = scipy.sparse.tril(self.Lc, k=-1).tocoo()
            keep = blk_of[coo.row] != blk_of[coo.col]
            self._Loff = scipy.sparse.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=(self.n, self.n))
            self._Loff.sort_indices()
        return self._Loff

    @property
    def LoffT(self):
        if self._LoffT is None:
            self._LoffT = self.Loff.T.tocsr()
            self._LoffT.sort_indices()
        return self._LoffT

    @property
    def nb(self):
        return len(self.bptr) - 1

    @property
    def H(self):
        return int(self.height.max()) if self.nb else 0


def factorize(S, cx, cy, leaf_size=8, n_parts=None, part_weights=None):
    """n_parts: make the dissection tree with that many equal subtrees (ordering.nested_dissection); the part of
    every block is kept in F.blk_part (-1: separators above the parts)."""
    n = S.shape[0]
    F = Factor()
    F.n = n
    if n_parts is None:
        perm, bptr, height, depth, dom = nested_dissection(S, cx, cy, leaf_size=leaf_size, return_tree=True)
        F.blk_part = None
    else:
        perm, bptr, height, depth, dom, F.blk_part = nested_dissection(S, cx, cy, leaf_size=leaf_size, return_tree=True,
                                                                      n_parts=n_parts, part_weights=part_weights)
    F.perm, F.bptr, F.height, F.depth, F.dom = perm, bptr, height, depth, dom
    Sp = scipy.sparse.csc_matrix(S)[perm][:, perm].tocsc()
    lu = scipy.sparse.linalg.splu(Sp, permc_spec="NATURAL", diag_pivot_thresh=0.0,
                                  options=dict(SymmetricMode=True))
    if not (np.array_equal(lu.perm_r, np.arange(n)) and np.array_equal(lu.perm_c, np.arange(n))):
        raise RuntimeError("factorisation pivoted; system matrix is not positive definite enough")
    d = lu.U.diagonal()
    if not np.all(d > 0):
        raise RuntimeError("system matrix is not positive definite")
    # Cholesky factor Lc = L diag(sqrt(d)): scale the columns of the unit lower factor in place, then one CSC -> CSR
    Lcsc = lu.L.tocsc()
    Lcsc.data *= np.repeat(np.sqrt(d), np.diff(Lcsc.indptr))
    Lc = Lcsc.tocsr()
    Lc.sort_indices()
    F.nnz_L = int(Lc.nnz)
    F.Lc = Lc
    F.Sp = Sp.tocsr()          # the permuted system matrix (subdomain.py recognises congruent subdomains by it)
    F.Sp.sort_indices()
    F.dinv = None              # inverses of the diagonal blocks: only the streaming program needs them (block_inverses)
    F._Loff = F._LoffT = None  # Lc without its diagonal blocks: built on first use (streaming program)
    return F


def block_inverses(F):
    """Explicit inverses of the diagonal blocks of the factor (computed on first use)."""
    if F.dinv is None:
        F.dinv = []
        for b in range(F.nb):
            r0, r1 = F.bptr[b], F.bptr[b + 1]
            D = F.Lc[r0:r1, r0:r1].toarray()
            F.dinv.append(scipy.linalg.solve_triangular(D, np.eye(r1 - r0), lower=True))
    return F.dinv


# ----------------------------------------------------------------------------------------------
# tiles
# ----------------------------------------------------------------------------------------------
def _pow2_floor(v):
    p = 1
    while p * 2 <= v:
        p *= 2
    return p


def _tile_rows_for(total_rows, want_tasks=16):
    return max(1, min(32, _pow2_floor(max(1, total_rows // want_tasks))))


def b_tiles(F, b, tr, transpose):
    """Phase-b tiles of block b (apply D^-1, or D^-T when transpose) in an order that is safe for
    in-place sequential execution. Returns a list of (row0, V, cols, flags)."""
    r0, r1 = int(F.bptr[b]), int(F.bptr[b + 1])
    k = r1 - r0
    Di = F.dinv[b].T if transpose else F.dinv[b]
    multi = k > tr
    starts = list(range(0, k, tr))
    if not transpose:
        starts.reverse()          # a lower-triangular row block reads rows above it: do the last rows first
    out = []
    for t0 in starts:
        t1 = min(k, t0 + tr)
        c0, c1 = (t0, k) if transpose else (0, t1)
        out.append((r0 + t0, Di[t0:t1, c0:c1], np.arange(r0 + c0, r0 + c1), TILE_STAGED if multi else 0))
    return out


def a_tiles(F, b, tr, M, col_mask=None):
    """Phase-a tiles of block b: out = src[row] - M[row, cols] src[cols] over columns outside the block.
    col_mask (bool per column) restricts the columns (resident engine: columns owned by one rank)."""
    r0, r1 = int(F.bptr[b]), int(F.bptr[b + 1])
    out = []
    for t0 in range(r0, r1, tr):
        t1 = min(r1, t0 + tr)
        sub = M[t0:t1]
        cols = np.unique(sub.indices)
        if col_mask is not None:
            cols = cols[col_mask[cols]]
        if cols.size == 0:
            continue
        out.append((t0, -sub[:, cols].toarray(), cols, TILE_SELF))
    return out


def pack_levels(levels, row_map=None, col_map=None):
    """
    levels: list of levels; a level is a list of groups; a group is a list of tiles (row0, V, cols, flags).
    row_map / col_map translate permuted row indices into the index space of the consumer.
    """
    levels = [[g for g in lev if g] for lev in levels]
    n_tiles = sum(len(g) for lev in levels for g in lev)
    n_groups = sum(len(lev) for lev in levels)
    level_ptr = np.zeros(len(levels) + 1, dtype=np.int32)
    group_ptr = np.zeros(n_groups + 1, dtype=np.int32)
    row0 = np.zeros(n_tiles, dtype=np.int32); nrows = np.zeros(n_tiles, dtype=np.int32)
    lpr = np.zeros(n_tiles, dtype=np.int32); nsteps = np.zeros(n_tiles, dtype=np.int32)
    flags = np.zeros(n_tiles, dtype=np.int32); stage_off = np.zeros(n_tiles, dtype=np.int32)
    col_off = np.zeros(n_tiles, dtype=np.int64); val_off = np.zeros(n_tiles, dtype=np.int64)
    cols_parts, vals_parts = [], []
    co = vo = t = gi = 0
    stage_rows = nnz = 0
    for li, lev in enumerate(levels):
        staged = 0
        for grp in lev:
            gi += 1
            for (r0, V, cols, fl) in grp:
                nr, nc = V.shape
                nrp = 1
                while nrp < nr:
                    nrp *= 2
                assert nrp <= 32
                m = 32 // nrp
                st = (nc + m - 1) // m
                Vp = np.zeros((nrp, st * m))
                Vp[:nr, :nc] = V
                cols = np.asarray(cols, dtype=np.int64)
                if col_map is not None:
                    cols = col_map[cols]
                    assert np.all(cols >= 0)
                cp = np.zeros(st * m, dtype=np.int32)
                cp[:nc] = cols
                if nc < st * m:
                    cp[nc:] = cols[-1]                  # padded columns: any valid index, value 0
                # [step][lane], lane = i*m + s, column = step*m + s
                packed = Vp.reshape(nrp, st, m).transpose(1, 0, 2).reshape(st, 32)
                row0[t] = r0 if row_map is None else row_map[r0]
                assert row0[t] >= 0
                nrows[t], lpr[t], nsteps[t], flags[t] = nr, m, st, fl
                col_off[t], val_off[t] = co, vo
                cols_parts.append(cp); vals_parts.append(packed.ravel())
                co += cp.size; vo += packed.size
                nnz += nr * nc
                if fl & TILE_STAGED:
                    stage_off[t] = staged
                    staged += nr
                t += 1
            group_ptr[gi] = t
        level_ptr[li + 1] = gi
        stage_rows = max(stage_rows, staged)
    return dict(level_ptr=level_ptr, group_ptr=group_ptr, tile_row0=row0, tile_nrows=nrows, tile_lpr=lpr,
                tile_nsteps=nsteps, tile_flags=flags, tile_stage_off=stage_off, tile_col_off=col_off,
                tile_val_off=val_off,
                cols=np.concatenate(cols_parts) if cols_parts else np.zeros(0, np.int32),
                vals=np.concatenate(vals_parts) if vals_parts else np.zeros(0),
                stage_rows=int(stage_rows), nnz=int(nnz))


RES_WARPS = 16          # warps per thread block of the subdomain kernel
STEP_BYTES = 320        # one stream step: 32 float64 values + 32 uint16 shared-memory element codes
CHAIN_ROWS = 32         # diagonal blocks up to this many rows are updated in place by a single warp


def _lpt(costs, n_warps):
    """Longest-processing-time assignment; returns lists of item indices per warp."""
    load = np.zeros(n_warps)
    assign = [[] for _ in range(n_warps)]
    for i in (np.argsort(-np.asarray(costs), kind="stable") if len(costs) else []):
        w = int(np.argmin(load))
        assign[w].append(int(i))
        load[w] += costs[i]
    return assign


class SolveProgram:
    """Streaming packing of the factor: perm (new-to-old faces) and sweeps['fwd'|'bwd'] (see pack_levels)."""

    def __init__(self):
        self.perm = None
        self.n = 0
        self.sweeps = {}
        self.stats = {}
        self.factor = None


def streaming_program(F):
    prog = SolveProgram()
    prog.n, prog.perm, prog.factor = F.n, F.perm.astype(np.int32), F
    block_inverses(F)
    H = F.H
    by_height = [np.flatnonzero(F.height == h) for h in range(H + 1)]
    rows_at = [int(sum(F.bptr[b + 1] - F.bptr[b] for b in by_height[h])) for h in range(H + 1)]
    fwd, bwd = [], []
    for h in range(H + 1):
        tr = _tile_rows_for(rows_at[h])
        if h > 0:
            fwd.append([[t] for b in by_height[h] for t in a_tiles(F, b, tr, F.Loff)])
        fwd.append([b_tiles(F, b, tr, False) for b in by_height[h]])
    for h in range(H, -1, -1):
        tr = _tile_rows_for(rows_at[h])
        if h < H:
            bwd.append([[t] for b in by_height[h] for t in a_tiles(F, b, tr, F.LoffT)])
        bwd.append([b_tiles(F, b, tr, True) for b in by_height[h]])
    prog.sweeps["fwd"] = pack_levels(fwd)
    prog.sweeps["bwd"] = pack_levels(bwd)
    prog.stats = dict(n=F.n, blocks=F.nb, height=H, nnz_L=F.nnz_L,
                      nnz_fwd=prog.sweeps["fwd"]["nnz"], nnz_bwd=prog.sweeps["bwd"]["nnz"],
                      vals_fwd=int(prog.sweeps["fwd"]["vals"].size), vals_bwd=int(prog.sweeps["bwd"]["vals"].size),
                      levels_fwd=len(fwd), levels_bwd=len(bwd))
    return prog


def build_solve_program(S, cx, cy, leaf_size=8):
    """factorize + streaming_program."""
    return streaming_program(factorize(S, cx, cy, leaf_size=leaf_size))


def apply_program_host(prog, b):
    """
    Host interpreter of the streaming solve program (float64 numpy) - used by CPU tests to validate the
    program itself against a direct solve; the device kernels implement exactly these semantics.
    b : (Nf, ...) right-hand side in ORIGINAL face numbering. Returns J in original numbering.
    """
    v = np.array(b, dtype=np.double)[prog.perm]
    for name in ("fwd", "bwd"):
        _run_packed(prog.sweeps[name], v)
    res = np.empty_like(v)
    res[prog.perm] = v
    return res


def _unpack_tile(sw, t):
    r0, nr, m, st = sw["tile_row0"][t], sw["tile_nrows"][t], sw["tile_lpr"][t], sw["tile_nsteps"][t]
    nrp = 32 // m
    cols = sw["cols"][sw["tile_col_off"][t]: sw["tile_col_off"][t] + st * m]
    packed = sw["vals"][sw["tile_val_off"][t]: sw["tile_val_off"][t] + st * 32]
    V = packed.reshape(st, nrp, m).transpose(1, 0, 2).reshape(nrp, st * m)[:nr]
    return r0, nr, cols, V


def _run_packed(sw, v, levels=None):
    gp = sw["group_ptr"]
    for l in (range(len(sw["level_ptr"]) - 1) if levels is None else levels):
        out = []
        for t in range(gp[sw["level_ptr"][l]], gp[sw["level_ptr"][l + 1]]):
            r0, nr, cols, V = _unpack_tile(sw, t)
            acc = np.tensordot(V, v[cols], axes=(1, 0))
            if sw["tile_flags"][t] & TILE_SELF:
                acc = acc + v[r0:r0 + nr]
            out.append((r0, nr, acc))
        for (r0, nr, acc) in out:
            v[r0:r0 + nr] = acc
