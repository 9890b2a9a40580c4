"""
Host-side factorisation of the cycle-space system and its compilation into a *solve program* for
the device.

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
"""
import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

from .ordering import nested_dissection

__all__ = ["SolveProgram", "build_solve_program", "system_matrix"]

TILE_SELF = 1        # add src[row] to the dot product (phase a)
TILE_STAGED = 2      # result must not be written before the whole level has been read (phase b, multi-tile block)


def system_matrix(A, Lmat, Rv, Cv):
    """A (L + diag(1/(Cv+Rv))) A^T as CSC float64. (reference: time_evolution.py:476,504-505)"""
    A = scipy.sparse.csc_matrix(A).astype(np.double)
    mid = scipy.sparse.csc_matrix(Lmat) + scipy.sparse.diags(1.0 / (Cv + Rv), 0)
    return scipy.sparse.csc_matrix(A @ mid @ A.T)


class SolveProgram:
    """
    Device-ready description of  J = S^-1 b  in the permuted face numbering.

    perm : (Nf,) new-to-old face permutation
    For each sweep ('fwd', 'bwd'):
      level_ptr : (n_levels + 1,) tiles of level l are level_ptr[l]:level_ptr[l+1]
      tile_row0, tile_nrows, tile_lpr (lanes per row), tile_nsteps, tile_flags : (n_tiles,) int32
      tile_col_off, tile_val_off : (n_tiles,) int64 offsets into cols / vals
      cols : int32 column indices, nsteps * lpr per tile
      vals : float64, nsteps * 32 per tile
      stage_rows : rows of staging needed (max over levels of rows in TILE_STAGED tiles)
    """

    def __init__(self):
        self.perm = None
        self.n = 0
        self.sweeps = {}
        self.stats = {}


def _pow2_floor(v):
    p = 1
    while p * 2 <= v:
        p *= 2
    return p


class _Emitter:
    def __init__(self):
        self.levels = []          # list of list of tiles
        self.cur = None

    def begin_level(self):
        self.cur = []
        self.levels.append(self.cur)

    def begin_group(self):
        """Tiles of one group are processed in the order given (streaming engine); groups of a level
        are independent."""
        self.cur.append([])

    def tile(self, row0, V, cols, flags):
        """V: (nrows_real, ncols) dense, rows row0..row0+nrows_real-1."""
        self.cur[-1].append((row0, V, np.asarray(cols, dtype=np.int32), flags))

    def pack(self, rows_per_tile_hint=None):
        self.levels = [[g for g in lev if g] for lev in self.levels]
        n_tiles = sum(len(g) for l in self.levels for g in l)
        n_groups = sum(len(l) for l in self.levels)
        level_ptr = np.zeros(len(self.levels) + 1, dtype=np.int32)
        group_ptr = np.zeros(n_groups + 1, dtype=np.int32)
        gi = 0
        row0 = np.zeros(n_tiles, dtype=np.int32); nrows = np.zeros(n_tiles, dtype=np.int32)
        lpr = np.zeros(n_tiles, dtype=np.int32); nsteps = np.zeros(n_tiles, dtype=np.int32)
        flags = np.zeros(n_tiles, dtype=np.int32)
        col_off = np.zeros(n_tiles, dtype=np.int64); val_off = np.zeros(n_tiles, dtype=np.int64)
        cols_parts, vals_parts = [], []
        co = vo = 0
        t = 0
        stage_rows = 0
        nnz = 0
        for li, lev in enumerate(self.levels):
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
                  cp = np.zeros(st * m, dtype=np.int32)
                  cp[:nc] = cols
                  if nc < st * m:
                      cp[nc:] = cols[-1] if nc else r0      # padded columns: any valid index, value 0
                  # [step][lane], lane = i*m + s, column = step*m + s
                  packed = Vp.reshape(nrp, st, m).transpose(1, 0, 2).reshape(st, 32)
                  row0[t], nrows[t], lpr[t], nsteps[t], flags[t] = r0, nr, m, st, fl
                  col_off[t], val_off[t] = co, vo
                  cols_parts.append(cp); vals_parts.append(packed.ravel())
                  co += cp.size; vo += packed.size
                  nnz += nr * nc
                  if fl & TILE_STAGED:
                      staged += nr
                  t += 1
              group_ptr[gi] = t
            level_ptr[li + 1] = gi
            stage_rows = max(stage_rows, staged)
        return dict(level_ptr=level_ptr, group_ptr=group_ptr, tile_row0=row0, tile_nrows=nrows, tile_lpr=lpr, tile_nsteps=nsteps,
                    tile_flags=flags, tile_col_off=col_off, tile_val_off=val_off,
                    cols=np.concatenate(cols_parts) if cols_parts else np.zeros(0, np.int32),
                    vals=np.concatenate(vals_parts) if vals_parts else np.zeros(0),
                    stage_rows=int(stage_rows), nnz=int(nnz))


def _tile_rows_for(total_rows, want_tasks=16):
    return max(1, min(32, _pow2_floor(max(1, total_rows // want_tasks))))


def build_solve_program(S, cx, cy, leaf_size=8, drop_tol=0.0):
    """
    S : (Nf, Nf) SPD scipy sparse matrix;  cx, cy : (Nf,) face centroids.
    Returns a SolveProgram.
    """
    n = S.shape[0]
    prog = SolveProgram()
    prog.n = n
    perm, bptr, height = nested_dissection(S, cx, cy, leaf_size=leaf_size)
    prog.perm = perm.astype(np.int32)
    Sp = scipy.sparse.csc_matrix(S)[perm][:, perm].tocsc()
    lu = scipy.sparse.linalg.splu(Sp, permc_spec="NATURAL", diag_pivot_thresh=0.0,
                                  options=dict(SymmetricMode=True))
    if not (np.array_equal(lu.perm_r, np.arange(n)) and np.array_equal(lu.perm_c, np.arange(n))):
        raise RuntimeError("factorisation pivoted; system matrix is not positive definite enough")
    d = lu.U.diagonal()
    if not np.all(d > 0):
        raise RuntimeError("system matrix is not positive definite")
    Lc = (lu.L @ scipy.sparse.diags(np.sqrt(d))).tocsr()       # Cholesky factor, S_perm = Lc Lc^T
    Lc.sort_indices()
    LcT = Lc.T.tocsr()                                          # rows of LcT = columns of Lc
    LcT.sort_indices()
    nb = len(bptr) - 1
    H = int(height.max()) if nb else 0
    by_height = [np.flatnonzero(height == h) for h in range(H + 1)]
    rows_at = [int(sum(bptr[b + 1] - bptr[b] for b in by_height[h])) for h in range(H + 1)]

    fwd, bwd = _Emitter(), _Emitter()
    dinv = {}
    for b in range(nb):
        r0, r1 = bptr[b], bptr[b + 1]
        D = Lc[r0:r1, r0:r1].toarray()
        dinv[b] = scipy.linalg.solve_triangular(D, np.eye(r1 - r0), lower=True)

    def emit_b_phase(em, h, transpose):
        em.begin_level()
        tr = _tile_rows_for(rows_at[h])
        for b in by_height[h]:
            r0, r1 = int(bptr[b]), int(bptr[b + 1])
            k = r1 - r0
            Di = dinv[b].T if transpose else dinv[b]
            multi = k > tr
            em.begin_group()
            starts = list(range(0, k, tr))
            if not transpose:
                starts.reverse()     # in-place safe order: a lower-triangular row block reads rows above it
            for t0 in starts:
                t1 = min(k, t0 + tr)
                if transpose:      # upper triangular: row i uses columns i..k-1
                    c0, c1 = t0, k
                else:              # lower triangular: row i uses columns 0..i
                    c0, c1 = 0, t1
                em.tile(r0 + t0, Di[t0:t1, c0:c1], np.arange(r0 + c0, r0 + c1), TILE_STAGED if multi else 0)

    def emit_a_phase(em, h, M):
        """M: CSR whose rows [r0:r1) hold the coupling coefficients (to be negated) outside the block."""
        em.begin_level()
        tr = _tile_rows_for(rows_at[h])
        for b in by_height[h]:
            r0, r1 = int(bptr[b]), int(bptr[b + 1])
            k = r1 - r0
            for t0 in range(0, k, tr):
                t1 = min(k, t0 + tr)
                sub = M[r0 + t0:r0 + t1]
                cols = np.unique(sub.indices)
                if cols.size == 0:
                    # still emit an identity tile? no: t_B = b_B unchanged, nothing to do
                    continue
                V = -sub[:, cols].toarray()
                em.begin_group()
                em.tile(r0 + t0, V, cols, TILE_SELF)

    # forward: heights ascending; a-phase uses Lc[B, <B]
    Lc_strict = scipy.sparse.tril(Lc, k=-1).tocsr()
    # remove within-block entries (they belong to D): keep only columns < block start
    blk_of = np.repeat(np.arange(nb), np.diff(bptr))
    coo = Lc_strict.tocoo()
    keep = blk_of[coo.row] != blk_of[coo.col]
    Loff = scipy.sparse.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=(n, n))
    LoffT = Loff.T.tocsr()      # row j holds Lc[i, j] for i in later blocks
    for h in range(H + 1):
        if h > 0:
            emit_a_phase(fwd, h, Loff)
        emit_b_phase(fwd, h, transpose=False)
    for h in range(H, -1, -1):
        if h < H:
            emit_a_phase(bwd, h, LoffT)
        emit_b_phase(bwd, h, transpose=True)
    prog.sweeps["fwd"] = fwd.pack()
    prog.sweeps["bwd"] = bwd.pack()
    prog.stats = dict(n=n, blocks=nb, height=H, nnz_L=int(Lc.nnz),
                      nnz_fwd=prog.sweeps["fwd"]["nnz"], nnz_bwd=prog.sweeps["bwd"]["nnz"],
                      vals_fwd=int(prog.sweeps["fwd"]["vals"].size), vals_bwd=int(prog.sweeps["bwd"]["vals"].size),
                      levels_fwd=len(fwd.levels), levels_bwd=len(bwd.levels))
    return prog


def apply_program_host(prog, b):
    """
    Host interpreter of the solve program (float64 numpy) - used by CPU tests to validate the program
    itself against a direct solve; the device kernels implement exactly these semantics.
    b : (Nf, ...) right-hand side in ORIGINAL face numbering. Returns J in original numbering.
    """
    v = np.array(b, dtype=np.double)[prog.perm]
    for name in ("fwd", "bwd"):
        sw = prog.sweeps[name]
        for l in range(len(sw["level_ptr"]) - 1):
            out = []
            gp = sw["group_ptr"]
            for t in range(gp[sw["level_ptr"][l]], gp[sw["level_ptr"][l + 1]]):
                r0, nr, m, st = sw["tile_row0"][t], sw["tile_nrows"][t], sw["tile_lpr"][t], sw["tile_nsteps"][t]
                nrp = 32 // m
                cols = sw["cols"][sw["tile_col_off"][t]: sw["tile_col_off"][t] + st * m]
                packed = sw["vals"][sw["tile_val_off"][t]: sw["tile_val_off"][t] + st * 32]
                V = packed.reshape(st, nrp, m).transpose(1, 0, 2).reshape(nrp, st * m)[:nr]
                acc = np.tensordot(V, v[cols], axes=(1, 0))
                if sw["tile_flags"][t] & TILE_SELF:
                    acc = acc + v[r0:r0 + nr]
                out.append((r0, nr, acc))
            for (r0, nr, acc) in out:
                v[r0:r0 + nr] = acc
    res = np.empty_like(v)
    res[prog.perm] = v
    return res
