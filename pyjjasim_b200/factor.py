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

    @property
    def nb(self):
        return len(self.bptr) - 1

    @property
    def H(self):
        return int(self.height.max()) if self.nb else 0


def factorize(S, cx, cy, leaf_size=8):
    n = S.shape[0]
    F = Factor()
    F.n = n
    perm, bptr, height, depth, dom = nested_dissection(S, cx, cy, leaf_size=leaf_size, return_tree=True)
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
STEP_BYTES = 320        # one stream step: 32 float64 values + 32 uint16 shared-memory rows
NO_ROW = 0xFFFF


def a_rows(F, blocks, M, col_mask=None):
    """Phase-a row tasks of the given blocks: (permuted row, cols, vals) with out = src[row] + sum vals*src[cols]."""
    out = []
    for b in blocks:
        for g in range(int(F.bptr[b]), int(F.bptr[b + 1])):
            lo, hi = M.indptr[g], M.indptr[g + 1]
            cols, vals = M.indices[lo:hi], -M.data[lo:hi]
            if col_mask is not None:
                keep = col_mask[cols]
                cols, vals = cols[keep], vals[keep]
            if cols.size:
                out.append((g, cols.astype(np.int64), vals))
    return out


def b_blocks(F, blocks, transpose):
    """Phase-b block tasks: (first permuted row, dense triangular matrix to apply to the block's own rows)."""
    return [(int(F.bptr[b]), F.dinv[b].T if transpose else F.dinv[b]) for b in blocks]


def _lpt(costs, n_warps):
    """Longest-processing-time assignment; returns (lists of item indices per warp, makespan)."""
    load = np.zeros(n_warps)
    assign = [[] for _ in range(n_warps)]
    for i in np.argsort(-np.asarray(costs), kind="stable") if len(costs) else []:
        w = int(np.argmin(load))
        assign[w].append(int(i))
        load[w] += costs[i]
    return assign, float(load.max()) if len(costs) else 0.0


def _tile_cost(steps, m):
    return steps + 3 + 2 * int(np.log2(m))


def _tiles_a(rows, m):
    """Group phase-a rows (sorted by length) into tiles of 32/m rows."""
    P = 32 // m
    order = sorted(range(len(rows)), key=lambda i: -rows[i][1].size)
    tiles = []
    for k in range(0, len(order), P):
        grp = [rows[i] for i in order[k:k + P]]
        tiles.append(dict(m=m, flags=TILE_SELF, rows=[(g, c, v) for (g, c, v) in grp]))
    return tiles


def _tiles_b(blocks, m, transpose):
    """Phase-b tiles: whole small blocks are packed together (updated in place by one warp), blocks with
    more rows than a tile holds are cut into row groups that write through the staging buffer."""
    P = 32 // m
    tiles = []
    small = sorted([b for b in blocks if b[1].shape[0] <= P], key=lambda b: -b[1].shape[0])
    bins = []
    for (r0, D) in small:
        k = D.shape[0]
        for bn in bins:
            if bn["free"] >= k:
                break
        else:
            bn = dict(free=P, rows=[])
            bins.append(bn)
        bn["free"] -= k
        for i in range(k):
            c0, c1 = (i, k) if transpose else (0, i + 1)
            bn["rows"].append((r0 + i, np.arange(r0 + c0, r0 + c1), D[i, c0:c1]))
    for bn in bins:
        tiles.append(dict(m=m, flags=0, rows=bn["rows"]))
    for (r0, D) in blocks:
        k = D.shape[0]
        if k <= P:
            continue
        for t0 in range(0, k, P):
            rows = []
            for i in range(t0, min(k, t0 + P)):
                c0, c1 = (i, k) if transpose else (0, i + 1)
                rows.append((r0 + i, np.arange(r0 + c0, r0 + c1), D[i, c0:c1]))
            tiles.append(dict(m=m, flags=TILE_STAGED, rows=rows))
    return tiles


def _tile_steps(t):
    m = t["m"]
    return 1 + max((r[1].size + m - 1) // m for r in t["rows"])


def plan_level(kind, items, transpose, n_warps):
    """Choose the lanes-per-row m that minimises the slowest warp's time and return (tiles, assignment)."""
    best = None
    for m in (1, 2, 4, 8, 16, 32):
        tiles = _tiles_a(items, m) if kind == "a" else _tiles_b(items, m, transpose)
        costs = [_tile_cost(_tile_steps(t), m) for t in tiles]
        assign, span = _lpt(costs, n_warps)
        if best is None or span < best[0]:
            best = (span, tiles, assign)
    return best[1], best[2]


def pack_ell_streams(level_specs, row_map, col_map, n_warps=RES_WARPS):
    """
    Resident-engine packing. level_specs: list of (kind 'a'|'b', items, transpose) per level, items from
    a_rows / b_blocks. Every (level, warp) pair gets a contiguous *stream* of 320-byte steps
    ([32 float64 values][32 uint16 shared-memory rows], one per lane) through which the warp prefetches
    straight across tile boundaries. A tile is 32/m output rows x m lanes per row; its first step carries the
    output row of every lane (0xFFFF: none), the following steps one (value, source row) pair per lane.
    Header: two int32 = (nrows-1) | log2(m) << 5 | flags << 8 , nsteps | stage_off << 16.
    """
    n_levels = len(level_specs)
    wt_ptr = np.zeros(n_levels * n_warps + 1, dtype=np.int32)
    ws_ptr = np.zeros(n_levels * n_warps + 1, dtype=np.int32)
    hdr, chunks = [], []
    n_steps = stage_rows = n_vals = 0
    staged_levels = []
    for li, (kind, items, transpose) in enumerate(level_specs):
        tiles, assign = plan_level(kind, items, transpose, n_warps) if items else ([], [[] for _ in range(n_warps)])
        staged = 0
        for t in tiles:
            if t["flags"] & TILE_STAGED:
                t["stage_off"] = staged
                staged += len(t["rows"])
        stage_rows = max(stage_rows, staged)
        staged_levels.append(int(staged))
        for w in range(n_warps):
            for ti in assign[w]:
                t = tiles[ti]
                m = t["m"]
                P = 32 // m
                st = _tile_steps(t)
                vals = np.zeros((st, 32))
                cols = np.zeros((st, 32), dtype=np.uint16)
                cols[0, :] = NO_ROW
                for i, (g, c, v) in enumerate(t["rows"]):
                    lanes = slice(i * m, (i + 1) * m)
                    out_row = int(row_map[g])
                    assert 0 <= out_row < NO_ROW
                    cols[0, lanes] = out_row
                    cm = col_map[np.asarray(c, dtype=np.int64)]
                    assert np.all(cm >= 0) and np.all(cm < NO_ROW)
                    n = cm.size
                    ns = (n + m - 1) // m
                    pc = np.zeros(ns * m, dtype=np.int64); pv = np.zeros(ns * m)
                    pc[:n] = cm; pv[:n] = v
                    pc[n:] = cm[-1]
                    cols[1:1 + ns, lanes] = pc.reshape(ns, m)
                    vals[1:1 + ns, lanes] = pv.reshape(ns, m)
                    n_vals += n
                rec = np.zeros((st, STEP_BYTES), dtype=np.uint8)
                rec[:, :256] = vals.view(np.uint8).reshape(st, 256)
                rec[:, 256:] = cols.view(np.uint8).reshape(st, 64)
                chunks.append(rec)
                assert st < 65536
                hdr.append(((len(t["rows"]) - 1) | (int(np.log2(m)) << 5) | (t["flags"] << 8),
                            st | (t.get("stage_off", 0) << 16)))
                n_steps += st
            wt_ptr[li * n_warps + w + 1] = len(hdr)
            ws_ptr[li * n_warps + w + 1] = n_steps
    stream = np.concatenate(chunks).ravel() if chunks else np.zeros(0, dtype=np.uint8)
    return dict(n_levels=n_levels, n_warps=n_warps, wt_ptr=wt_ptr, ws_ptr=ws_ptr,
                thdr=np.asarray(hdr, dtype=np.int32).reshape(-1, 2), stream=stream, n_steps=n_steps,
                stage_rows=int(stage_rows), vals=int(n_vals), staged_rows=staged_levels)


def _run_stream_level(ps, v, level):
    """Host interpreter of one level of a warp-stream program (mirrors the device kernel)."""
    nw = ps["n_warps"]
    out = []
    rec = ps["stream"].reshape(-1, STEP_BYTES)
    for w in range(nw):
        idx = level * nw + w
        s = ps["ws_ptr"][idx]
        for t in range(ps["wt_ptr"][idx], ps["wt_ptr"][idx + 1]):
            h0, h1 = int(ps["thdr"][t, 0]), int(ps["thdr"][t, 1])
            nr, mshift, fl = (h0 & 31) + 1, (h0 >> 5) & 7, (h0 >> 8) & 3
            st = h1 & 0xffff
            m = 1 << mshift
            vals = rec[s:s + st, :256].copy().view(np.float64).reshape(st, 32)
            cols = rec[s:s + st, 256:].copy().view(np.uint16).reshape(st, 32).astype(np.int64)
            s += st
            rows = cols[0, ::m][:nr]
            prod = vals[1:][(...,) + (None,) * (v.ndim - 1)] * v[cols[1:]]    # (st-1, 32, ...)
            lane_sum = prod.sum(axis=0)
            acc = lane_sum.reshape((32 // m, m) + v.shape[1:]).sum(axis=1)[:nr]
            if fl & TILE_SELF:
                acc = acc + v[rows]
            out.append((rows, acc))
        assert s == ps["ws_ptr"][idx + 1]
    for (rows, acc) in out:
        v[rows] = acc


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
        ps = pack_ell_streams(specs, row_map=smem_index[r], col_map=smem_index[r])
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
