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

Two packings of the same tiles exist:
  * ``streaming_program``  - rows are permuted face indices (streaming engine, any size);
  * ``resident_plan``      - the elimination tree is cut at depth log2(C); each of the C thread blocks
    of a cluster owns one subtree (its rows live in that block's shared memory), the separators above
    the cut are replicated in every block and combined with an all-reduce over distributed shared
    memory in the forward sweep (resident engine).
"""
import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

from .ordering import nested_dissection

__all__ = ["Factor", "factorize", "streaming_program", "resident_plan", "system_matrix", "apply_program_host",
           "build_solve_program"]

TILE_SELF = 1        # add src[row] to the dot product (phase a)
TILE_STAGED = 2      # block spans several tiles: results must not overwrite src before the level is complete

OP_LEVEL, OP_ALLREDUCE = 0, 1


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
        self.Loff = None          # CSR: Lc without the diagonal blocks  (row i: couplings to earlier blocks)
        self.LoffT = None         # CSR: its transpose                   (row j: couplings to later blocks)
        self.nnz_L = 0
        self.blk_part = None      # part (subtree) of each block when the ordering was made with n_parts
        self.Lc = None            # CSR: the full Cholesky factor (subdomain engine: Schur complement of the top separators)

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
    Lc = (lu.L @ scipy.sparse.diags(np.sqrt(d))).tocsr()       # Cholesky factor
    Lc.sort_indices()
    F.nnz_L = int(Lc.nnz)
    F.Lc = Lc
    nb = F.nb
    F.dinv = []
    for b in range(nb):
        r0, r1 = bptr[b], bptr[b + 1]
        D = Lc[r0:r1, r0:r1].toarray()
        F.dinv.append(scipy.linalg.solve_triangular(D, np.eye(r1 - r0), lower=True))
    blk_of = np.repeat(np.arange(nb), np.diff(bptr))
    coo = scipy.sparse.tril(Lc, k=-1).tocoo()
    keep = blk_of[coo.row] != blk_of[coo.col]
    F.Loff = scipy.sparse.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=(n, n))
    F.Loff.sort_indices()
    F.LoffT = F.Loff.T.tocsr()
    F.LoffT.sort_indices()
    return F


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


RES_WARPS = 16          # warps per thread block of the resident kernel
RES_WT = 8              # problems per tile of the resident engine (= N of the m8n8k4 FP64 MMA)
STEP_BYTES = 320        # one stream step: 32 float64 values + 32 uint16 shared-memory element codes
CHAIN_ROWS = 32         # diagonal blocks up to this many rows are updated in place by a single warp


def a_rows(F, blocks, M, col_mask=None):
    """Phase-a tasks of the given blocks: per block (first permuted row, CSR rows of M restricted to col_mask)."""
    out = []
    for b in blocks:
        r0, r1 = int(F.bptr[b]), int(F.bptr[b + 1])
        out.append((r0, r1, M, col_mask))
    return out


def b_blocks(F, blocks, transpose):
    """Phase-b block tasks: (first permuted row, dense triangular matrix to apply to the block's own rows)."""
    return [(int(F.bptr[b]), F.dinv[b].T if transpose else F.dinv[b]) for b in blocks]


def _lpt(costs, n_warps):
    """Longest-processing-time assignment; returns lists of item indices per warp."""
    load = np.zeros(n_warps)
    assign = [[] for _ in range(n_warps)]
    for i in (np.argsort(-np.asarray(costs), kind="stable") if len(costs) else []):
        w = int(np.argmin(load))
        assign[w].append(int(i))
        load[w] += costs[i]
    return assign


def _mma_tiles_a(items):
    """8-row tiles of phase a: out = src[row] - M[row, cols] src[cols]; one unit per tile."""
    units = []
    for (r0, r1, M, col_mask) in items:
        for t0 in range(r0, r1, 8):
            t1 = min(r1, t0 + 8)
            sub = M[t0:t1]
            cols = np.unique(sub.indices)
            if col_mask is not None:
                cols = cols[col_mask[cols]]
            if cols.size == 0:
                continue
            units.append([dict(row0=t0, V=-sub[:, cols].toarray(), cols=cols.astype(np.int64), flags=TILE_SELF)])
    return units


def _mma_tiles_b(blocks, transpose):
    """8-row tiles of phase b. A block of at most CHAIN_ROWS rows is one unit: its tiles are processed by one
    warp in an order that makes the in-place update safe; larger blocks write through the staging buffer."""
    units = []
    for (r0, D) in blocks:
        k = D.shape[0]
        starts = list(range(0, k, 8))
        if not transpose:
            starts.reverse()          # lower triangular: a row group reads the rows above it -> last group first
        tiles = []
        for t0 in starts:
            t1 = min(k, t0 + 8)
            c0, c1 = (t0, k) if transpose else (0, t1)
            tiles.append(dict(row0=r0 + t0, V=D[t0:t1, c0:c1], cols=np.arange(r0 + c0, r0 + c1, dtype=np.int64), flags=0))
        if k <= CHAIN_ROWS:
            units.append(tiles)
        else:
            for t in tiles:
                t["flags"] = TILE_STAGED
                units.append([t])
    return units


def smem_code(col):
    """Element offset (in float64) of problem n = 0..7 of shared-memory row `col` in the swizzled layout:
    16-byte chunk c of row r lives at chunk position c ^ ((r >> 1) & 3)."""
    col = np.asarray(col, dtype=np.int64)[..., None]
    n = np.arange(8)
    return col * 8 + (((n >> 1) ^ ((col >> 1) & 3)) << 1) + (n & 1)


def pack_mma_streams(level_specs, row_map, col_map, n_warps=RES_WARPS):
    """
    Resident-engine packing for the FP64 tensor-core sweep. level_specs: list of (kind 'a'|'b', items,
    transpose) per level. A tile is 8 output rows x K columns (K a multiple of 4) of a dense matrix V; the
    device computes C(8 rows x 8 problems) += V(8x4) . S(4x8) per stream step with mma.m8n8k4, S gathered from
    the shared-memory vector. Every (level, warp) pair owns a contiguous stream of 320-byte steps:
      32 float64: lane = row*4 + kk holds V[row, 4*step + kk]          (the A fragment)
      32 uint16 : lane = n*4 + kk holds the element code of src[col(4*step + kk)], problem n   (the B fragment)
    Header (two int32): row0 | (nrows-1) << 16 | flags << 19 ;  nsteps | stage_off << 16.
    """
    n_levels = len(level_specs)
    wt_ptr = np.zeros(n_levels * n_warps + 1, dtype=np.int32)
    ws_ptr = np.zeros(n_levels * n_warps + 1, dtype=np.int32)
    hdr, chunks = [], []
    n_steps = stage_rows = n_vals = 0
    staged_rows = []
    lane = np.arange(32)
    for li, (kind, items, transpose) in enumerate(level_specs):
        units = (_mma_tiles_a(items) if kind == "a" else _mma_tiles_b(items, transpose)) if items else []
        costs = [sum((t["V"].shape[1] + 3) // 4 + 5 for t in u) for u in units]
        assign = _lpt(costs, n_warps)
        staged = 0
        for u in units:
            for t in u:
                if t["flags"] & TILE_STAGED:
                    t["stage_off"] = staged
                    staged += t["V"].shape[0]
        stage_rows = max(stage_rows, staged)
        staged_rows.append(int(staged))
        for w in range(n_warps):
            for ui in assign[w]:
                for t in units[ui]:
                    V = t["V"]
                    nr, nc = V.shape
                    st = (nc + 3) // 4
                    Vp = np.zeros((8, st * 4))
                    Vp[:nr, :nc] = V
                    cm = col_map[t["cols"]]
                    assert np.all(cm >= 0) and np.all(cm < 8192)
                    cp = np.full(st * 4, cm[-1], dtype=np.int64)
                    cp[:nc] = cm
                    vals = Vp.reshape(8, st, 4).transpose(1, 0, 2).reshape(st, 32)       # lane = row*4 + kk
                    codes = smem_code(cp.reshape(st, 4))                                  # (st, 4, 8): [kk][n]
                    codes = codes.transpose(0, 2, 1).reshape(st, 32).astype(np.uint16)    # lane = n*4 + kk
                    rec = np.zeros((st, STEP_BYTES), dtype=np.uint8)
                    rec[:, :256] = np.ascontiguousarray(vals).view(np.uint8).reshape(st, 256)
                    rec[:, 256:] = np.ascontiguousarray(codes).view(np.uint8).reshape(st, 64)
                    chunks.append(rec)
                    row = int(row_map[t["row0"]])
                    assert 0 <= row < 65536 and st < 65536
                    hdr.append((row | ((nr - 1) << 16) | (t["flags"] << 19), st | (t.get("stage_off", 0) << 16)))
                    n_steps += st
                    n_vals += nr * nc
            wt_ptr[li * n_warps + w + 1] = len(hdr)
            ws_ptr[li * n_warps + w + 1] = n_steps
    stream = np.concatenate(chunks).ravel() if chunks else np.zeros(0, dtype=np.uint8)
    return dict(n_levels=n_levels, n_warps=n_warps, wt_ptr=wt_ptr, ws_ptr=ws_ptr,
                thdr=np.asarray(hdr, dtype=np.int32).reshape(-1, 2), stream=stream, n_steps=n_steps,
                stage_rows=int(stage_rows), vals=int(n_vals), staged_rows=staged_rows)


def _run_stream_level(ps, v, level):
    """Host interpreter of one level of an MMA stream program (mirrors the device kernel, including the
    sequential in-place semantics of the tiles of one warp). v: (rows, 8) float64."""
    assert v.ndim == 2 and v.shape[1] == 8
    nw = ps["n_warps"]
    rec = ps["stream"].reshape(-1, STEP_BYTES)
    flat = v.reshape(-1)
    # swizzled view of v: the codes address the physical layout, so build it once per tile
    def physical(vv):
        rows = np.arange(vv.shape[0])
        ph = np.empty_like(vv).reshape(vv.shape[0], 4, 2)
        lg = vv.reshape(vv.shape[0], 4, 2)
        for c in range(4):
            ph[rows, c ^ ((rows >> 1) & 3)] = lg[rows, c]
        return ph.reshape(-1)
    staged = []
    for w in range(nw):
        idx = level * nw + w
        s = ps["ws_ptr"][idx]
        for t in range(ps["wt_ptr"][idx], ps["wt_ptr"][idx + 1]):
            h0, h1 = int(ps["thdr"][t, 0]), int(ps["thdr"][t, 1])
            row0, nr, fl = h0 & 0xffff, ((h0 >> 16) & 7) + 1, (h0 >> 19) & 3
            st = h1 & 0xffff
            vals = rec[s:s + st, :256].copy().view(np.float64).reshape(st, 8, 4)          # [step][row][kk]
            codes = rec[s:s + st, 256:].copy().view(np.uint16).reshape(st, 8, 4).astype(np.int64)   # [step][n][kk]
            s += st
            ph = physical(v)
            S = ph[codes]                                                                  # [step][n][kk]
            acc = np.einsum("srk,snk->rn", vals, S)[:nr]
            if fl & TILE_SELF:
                acc = acc + v[row0:row0 + nr]
            if fl & TILE_STAGED:
                staged.append((row0, nr, acc))
            else:
                v[row0:row0 + nr] = acc            # in place, in the warp's program order
        assert s == ps["ws_ptr"][idx + 1]
    for (row0, nr, acc) in staged:
        v[row0:row0 + nr] = acc


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


# ----------------------------------------------------------------------------------------------
# resident plan
# ----------------------------------------------------------------------------------------------
class ResidentPlan:
    """
    Per-rank programs for a cluster of C thread blocks (see module docstring).

    C, n_rows            : cluster size, rows of each block's shared-memory vector (local + replicated)
    n_local_max, n_shared
    smem_index[r]        : (Nf,) shared-memory row of permuted face g on rank r, or -1
    owner_rank           : (Nf,) rank whose vector holds the authoritative copy of face g (-1: replicated)
    prog[r]              : pack_levels(...) result in rank r's index space; one level per OP_LEVEL op
    ops                  : (n_ops, 4) int32, identical structure on every rank:
                           (OP_LEVEL, level index, staged?, 0) or (OP_ALLREDUCE, row_lo, row_hi, 0)
    n_fwd_ops            : ops[:n_fwd_ops] form the forward sweep, the rest the backward sweep
    """


def resident_plan(F, C, want_tasks=16):
    assert C in (1, 2, 4, 8, 16)
    d = int(np.log2(C))
    nb, n = F.nb, F.n
    shared_blk = F.depth < d
    blk_rank = np.where(shared_blk, -1, F.dom >> np.maximum(F.depth - d, 0))
    sizes = np.diff(F.bptr)
    blk_of = np.repeat(np.arange(nb), sizes)
    row_rank = blk_rank[blk_of]                                   # -1: replicated
    n_local = np.array([int(np.sum(row_rank == r)) for r in range(C)])
    n_local_max = int(n_local.max()) if C else 0
    # replicated rows: ordered by (height, row) so that one all-reduce covers a contiguous range
    sh_blocks = np.flatnonzero(shared_blk)
    sh_blocks = sh_blocks[np.lexsort((sh_blocks, F.height[sh_blocks]))]
    sh_index = np.full(n, -1, dtype=np.int64)
    pos = 0
    sh_range = {}
    for b in sh_blocks:
        k = int(sizes[b])
        sh_index[F.bptr[b]:F.bptr[b + 1]] = n_local_max + pos + np.arange(k)
        h = int(F.height[b])
        lo, hi = sh_range.get(h, (n_local_max + pos, n_local_max + pos))
        sh_range[h] = (lo, n_local_max + pos + k)
        pos += k
    n_shared = pos
    smem_index = []
    for r in range(C):
        m = sh_index.copy()
        rows = np.flatnonzero(row_rank == r)
        m[rows] = np.arange(rows.size)
        smem_index.append(m)
    # columns of replicated blocks are contributed by the lowest rank below the block
    col_owner = np.where(row_rank >= 0, row_rank,
                         (F.dom[blk_of] << np.maximum(d - F.depth[blk_of], 0)))
    col_owner = np.minimum(col_owner, C - 1)

    local_blocks = [np.flatnonzero(blk_rank == r) for r in range(C)]
    Hloc = max([int(F.height[lb].max()) if lb.size else 0 for lb in local_blocks] + [0])
    sh_heights = sorted(sh_range)

    def tr_for(blocks):
        return _tile_rows_for(int(sum(sizes[b] for b in blocks)), want_tasks)

    if n_local_max + n_shared > 8192:
        raise ValueError("resident plan: %d rows per block exceed the 16-bit element codes" % (n_local_max + n_shared))
    plan = ResidentPlan()
    plan.C, plan.n_local_max, plan.n_shared, plan.n_rows = C, n_local_max, n_shared, n_local_max + n_shared
    plan.smem_index, plan.row_rank, plan.col_owner = smem_index, row_rank, col_owner
    plan.prog, ops = [], None
    for r in range(C):
        specs, rops = [], []

        def level(kind, items, transpose=False):
            specs.append((kind, items, transpose))
            rops.append([OP_LEVEL, len(specs) - 1, 0, 0])

        # ---- forward: local subtree bottom-up
        for h in range(Hloc + 1):
            blocks = [b for b in local_blocks[r] if F.height[b] == h]
            if h > 0:
                level("a", a_rows(F, blocks, F.Loff))
            level("b", b_blocks(F, blocks, False), False)
        # ---- forward: replicated separators bottom-up, partial sums + all-reduce
        for h in sh_heights:
            blocks = [b for b in sh_blocks if F.height[b] == h]
            level("a", a_rows(F, blocks, F.Loff, col_mask=(col_owner == r)))
            rops.append([OP_ALLREDUCE, sh_range[h][0], sh_range[h][1], 0])
            level("b", b_blocks(F, blocks, False), False)
        n_fwd = len(rops)
        # ---- backward: replicated separators top-down (computed redundantly by every rank)
        for h in reversed(sh_heights):
            blocks = [b for b in sh_blocks if F.height[b] == h]
            level("a", a_rows(F, blocks, F.LoffT))
            level("b", b_blocks(F, blocks, True), True)
        # ---- backward: local subtree top-down
        for h in range(Hloc, -1, -1):
            blocks = [b for b in local_blocks[r] if F.height[b] == h]
            level("a", a_rows(F, blocks, F.LoffT))
            level("b", b_blocks(F, blocks, True), True)
        ps = pack_mma_streams(specs, row_map=smem_index[r], col_map=smem_index[r])
        plan.prog.append(ps)
        for op in rops:
            if op[0] == OP_LEVEL:
                op[2] = ps["staged_rows"][op[1]]        # number of staged rows of the level (0: none)
        rops = np.array(rops, dtype=np.int32).reshape(-1, 4)
        if ops is None:
            ops, plan.n_fwd_ops = rops, n_fwd
        else:
            # identical structure on every rank; the staged-row count of a level is per rank
            assert np.array_equal(ops[:, [0, 1]], rops[:, [0, 1]])
            assert np.array_equal(ops[ops[:, 0] == OP_ALLREDUCE], rops[rops[:, 0] == OP_ALLREDUCE])
        plan.rank_ops = getattr(plan, "rank_ops", []) + [rops]
    plan.ops = np.stack(plan.rank_ops)            # (C, n_ops, 4)
    plan.stage_rows = max(p["stage_rows"] for p in plan.prog)
    plan.allreduce_rows = max([hi - lo for (lo, hi) in sh_range.values()] + [0])
    plan.vals = [int(p["vals"]) for p in plan.prog]
    return plan


def apply_resident_plan_host(F, plan, b_perm):
    """Host interpreter of a resident plan (CPU test of the plan; mirrors the device kernel's data flow).
    b_perm: (Nf, ...) right-hand side in PERMUTED numbering. Returns J in permuted numbering."""
    C = plan.C
    shape_tail = b_perm.shape[1:]
    vec = [np.zeros((plan.n_rows,) + shape_tail) for _ in range(C)]
    # every rank holds its local rows; replicated rows start as PARTIAL right-hand sides: put everything on rank 0
    for r in range(C):
        loc = np.flatnonzero(plan.row_rank == r)
        vec[r][plan.smem_index[r][loc]] = b_perm[loc]
    sh = np.flatnonzero(plan.row_rank < 0)
    if sh.size:
        # split the replicated right-hand side unevenly over the ranks to exercise the reduction
        w = np.linspace(1.0, 2.0, C)
        w = w / w.sum()
        for r in range(C):
            vec[r][plan.smem_index[r][sh]] = w[r] * b_perm[sh]
    for op in plan.ops[0]:
        if op[0] == OP_LEVEL:
            for r in range(C):
                _run_stream_level(plan.prog[r], vec[r], int(op[1]))
        else:
            lo, hi = op[1], op[2]
            tot = sum(vec[r][lo:hi] for r in range(C))
            for r in range(C):
                vec[r][lo:hi] = tot
    out = np.zeros_like(b_perm)
    for r in range(C):
        loc = np.flatnonzero(plan.row_rank == r)
        out[loc] = vec[r][plan.smem_index[r][loc]]
    if sh.size:
        out[sh] = vec[0][plan.smem_index[0][sh]]
        for r in range(1, C):
            assert np.allclose(vec[r][plan.smem_index[r][sh]], out[sh], rtol=1e-12, atol=1e-13)
    return out
