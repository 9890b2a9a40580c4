"""
Host-side plan of the SUBDOMAIN step engine (device side: csrc/jj_subdomain.cu).

What it replaces: the two SuperLU triangular sweeps per time step of the reference
(reference: time_evolution.py:506, :560-569) and the projections A., A^T. around them (:560-570).

The block elimination tree of the nested dissection (ordering.py, factor.py) is cut at depth d:

  * the P = 2^d subtrees below the cut are SUBDOMAINS. They are mutually uncoupled, so a thread block
    owns one (subdomain s, chunk of PC = 8*NG problems) pair and does everything that is local to it
    out of its shared memory: backward substitution of the local rows, the junction update, the face
    projection and the forward elimination of the next step. The local part of the factor is streamed
    from L2 as FP64 tensor-core fragments and reused by all PC problems.
  * the separators above the cut are the TOP rows (n_top of them). Eliminating every subdomain leaves
    the Schur complement S_top = L_TT L_TT^T on them; the device applies its explicit inverse as ONE dense
    (n_top x n_top) x (n_top x W) FP64 tensor-core product per time step, spread over all SMs, instead
    of log2(P) more dependent tree levels.

Per time step the device therefore runs   [local: bwd, junctions, faces, fwd] -> grid barrier ->
[assemble r_top] -> grid barrier -> [J_top = S_top^-1 r_top] -> grid barrier.

Data exchanged through global memory (L2 resident): per (chunk, subdomain) the contribution of the
subdomain to the top right-hand side (`halo` rows = top rows coupled to the subdomain), and J_top.

A subdomain's shared-memory vector has rows [0, n_loc) = its local faces (ascending permuted index) and
rows [n_loc, n_loc + n_halo) = its halo rows (ascending top index); a row holds PC float64 as 32-byte chunks
of 4 problems, chunk c of row r is stored at chunk position c ^ (r & 3) (bank-conflict-free tensor-core gathers).
"""
import hashlib
import os

import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.csgraph

from .factor import TILE_SELF, TILE_STAGED, CHAIN_ROWS, STEP_BYTES, RES_WARPS, _lpt

__all__ = ["SubdomainPlan", "subdomain_plan", "apply_subdomain_plan_host", "elem_code"]

RING_STEPS = 4          # the device walks a tile's stream in ring blocks of this many steps (csrc/jj_subdomain.cu: RING)


def elem_code(row, NG):
    """Element offset (float64 units) of problem n = 0..7 of group 0 of shared-memory row `row`: a row is PC/4
    chunks of 4 float64 and chunk c sits at position c ^ (row & 3) (c ^ (row & 1) when PC = 8);
    group g lives at ``code ^ (g << 3)``."""
    row = np.asarray(row, dtype=np.int64)[..., None]
    n = np.arange(8)
    cm = 3 if NG >= 2 else 1
    return row * (8 * NG) + (((n >> 2) ^ (row & cm)) << 2) + (n & 3)


class SubdomainPlan:
    """See module docstring. Attributes are plain numpy arrays ready for the C ABI (JJSubdomainPlan)."""


def _tile_steps(nc):
    """stream steps of a tile with nc columns: whole ring blocks (zero steps at the end)"""
    return ((nc + 3) // 4 + RING_STEPS - 1) // RING_STEPS * RING_STEPS


def _pack_levels(levels, NG, n_warps, n_bwd=0):
    """levels: list of unit lists; a unit is a list of tiles dict(row0, V, cols, flags) executed in order by
    one warp. Returns the per-(level, warp) tile ranges and streams. Units longer than a warp's fair share of a
    level (and units of levels with fewer units than warps) are split over problem groups: a warp then handles
    ng < NG groups of the unit; the copies re-read the unit's (small) stream but run concurrently.
    Tiles and stream steps are emitted warp-major inside each sweep (levels [0, n_bwd) and [n_bwd, ...)), so the
    stream of a warp is contiguous across the levels of a sweep and its prefetch ring never drains.
    A stream step is 320 bytes: the A fragment (32 float64, lane = row*4 + kk) and the shared-memory element codes of
    the B fragment (32 uint16, lane = n*4 + kk)."""
    n_levels = len(levels)
    wt_ptr = np.zeros((n_levels * n_warps, 2), dtype=np.int32)
    ws_ptr = np.zeros(n_levels * n_warps, dtype=np.int32)
    hdr, emitted, lstaged = [], [], []
    n_steps = n_vals = 0
    assigned = []                      # per level: per warp list of (unit index, g0, ng)
    for li, units in enumerate(levels):
        staged = 0
        for u in units:
            for t in u:
                t["st"] = _tile_steps(t["V"].shape[1])
                t["stage_off"] = 0
                if t["flags"] & TILE_STAGED:
                    t["stage_off"] = staged
                    staged += t["V"].shape[0]
        lstaged.append(staged)
        ucost = [sum(t["st"] for t in u) for u in units]
        split = [1] * len(units)
        while True:
            share = sum(c * NG + 6 * sp for c, sp in zip(ucost, split)) / float(n_warps)
            worst = max(range(len(units)), key=lambda i: ucost[i] * (NG // split[i])) if units else None
            if worst is None or split[worst] >= NG:
                break
            if ucost[worst] * (NG // split[worst]) > 1.25 * share + 4 or sum(split) * 2 <= n_warps:
                split[worst] *= 2
            else:
                break
        tasks = [(ui, g0, NG // split[ui]) for ui in range(len(units)) for g0 in range(0, NG, NG // split[ui])]
        costs = [ucost[ui] * ng + 6 * len(units[ui]) for (ui, g0, ng) in tasks]
        assign = _lpt(costs, n_warps)
        assigned.append([[tasks[ti] for ti in assign[w]] for w in range(n_warps)])
    for (l0, l1) in ((0, n_bwd), (n_bwd, n_levels)):
        for w in range(n_warps):
            for li in range(l0, l1):
                wt_ptr[li * n_warps + w, 0] = len(hdr)
                ws_ptr[li * n_warps + w] = n_steps
                for (ui, g0, ng) in assigned[li][w]:
                    for t in levels[li][ui]:
                        nr, nc = t["V"].shape
                        st = t["st"]
                        assert 0 <= t["row0"] < 65536 and st < 65536 and t["stage_off"] < 32768
                        hdr.append((t["row0"] | ((nr - 1) << 16) | (t["flags"] << 19) | (g0 << 21) | ((ng - 1) << 25),
                                    st | (t["stage_off"] << 16)))
                        emitted.append((n_steps, t))
                        n_steps += st
                        if g0 == 0:
                            n_vals += nr * nc
                wt_ptr[li * n_warps + w, 1] = len(hdr)
    # all records at once: values as (8 rows, all columns), column rows as one list, one transposition at the end
    VA = np.zeros((8, n_steps * 4))
    CA = np.zeros(n_steps * 4, dtype=np.int64)
    for (s0, t) in emitted:
        nr, nc = t["V"].shape
        VA[:nr, 4 * s0:4 * s0 + nc] = t["V"]
        CA[4 * s0:4 * s0 + nc] = t["cols"]
        CA[4 * s0 + nc:4 * (s0 + t["st"])] = t["cols"][-1]
    stream = np.zeros((n_steps, STEP_BYTES), dtype=np.uint8)
    if n_steps:
        vals = np.ascontiguousarray(VA.reshape(8, n_steps, 4).transpose(1, 0, 2))            # [step][row][kk]
        codes = elem_code(CA.reshape(n_steps, 4), NG)                                        # [step][kk][n]
        assert codes.max() < 65536
        codes = np.ascontiguousarray(codes.transpose(0, 2, 1)).astype(np.uint16)             # [step][n][kk]
        stream[:, :256] = vals.view(np.uint8).reshape(n_steps, 256)
        stream[:, 256:] = codes.view(np.uint8).reshape(n_steps, 64)
    return dict(n_levels=n_levels, n_warps=n_warps, wt_ptr=wt_ptr, ws_ptr=ws_ptr,
                thdr=np.asarray(hdr, dtype=np.int32).reshape(-1, 2), stream=stream.ravel(), n_steps=int(n_steps),
                lstaged=np.asarray(lstaged, dtype=np.int32), stage_rows=int(max(lstaged) if lstaged else 0),
                vals=int(n_vals))


def _background(fn, *args):
    """start fn(*args) on a thread; the returned callable waits for it and returns its result (or re-raises)"""
    import threading
    box = {}

    def run():
        try:
            box["value"] = fn(*args)
        except BaseException as e:        # handed to the caller
            box["error"] = e
    t = threading.Thread(target=run, daemon=True)
    t.start()

    def result():
        t.join()
        if "error" in box:
            raise box["error"]
        return box["value"]
    return result


def _tri_inverse(L):
    """inverse of a dense lower-triangular matrix (LAPACK dtrtri)"""
    inv, info = scipy.linalg.lapack.dtrtri(np.asfortranarray(L), lower=1)
    if info != 0:
        raise np.linalg.LinAlgError("singular diagonal block in the factor")
    return np.tril(inv)


def _tiles_dense(M, row0, col_index, flags=TILE_SELF):
    """8-row tiles of out[row0 + i] (+)= -M[i, :] src: M is a dense array, col_index maps its columns to vector rows."""
    units = []
    nz = M != 0
    for t0 in range(0, M.shape[0], 8):
        cols = np.flatnonzero(nz[t0:t0 + 8].any(axis=0))
        if cols.size == 0:
            continue
        units.append([dict(row0=int(row0 + t0), V=-M[t0:t0 + 8, cols], cols=col_index[cols], flags=flags)])
    return units


def _subdomain_inputs(F, loc, hrows, blk_of):
    """What the sweep program of a subdomain depends on: tree heights of its rows, its block of the factor (CSR) and
    the factor rows of its halo (dense, local columns in the order of loc), plus a digest of all three. Two subdomains
    with equal digests (the translated copies of a regular lattice) share one program - on the host and on the
    device, where the shared stream stays in L2."""
    n = loc.size
    hrow = F.height[blk_of[loc]].astype(np.int64)
    lo, hi = int(loc[0]), int(loc[-1]) + 1
    contiguous = hi - lo == n            # a subtree of the dissection is a contiguous range of the post-order
    L0 = scipy.sparse.csr_matrix(F.Lc[lo:hi, lo:hi] if contiguous else F.Lc[loc][:, loc])
    if hrows.size:
        Lh = F.Lc[hrows]                 # separator rows: their local columns are off-block entries of the factor
        Lh = (Lh[:, lo:hi] if contiguous else Lh[:, loc]).toarray()
    else:
        Lh = np.zeros((0, n))
    # digest of the SYSTEM matrix blocks that determine (L0, Lh) - L0 is the Cholesky factor of S[loc, loc] and
    # Lh = S[halo, loc] L0^-T, a leaf subtree being eliminated before everything it touches - and of the row heights.
    # The factor values themselves differ in the last bit between congruent subdomains (summation orders inside the
    # supernodal factorisation), the matrix entries do not.
    S0 = F.Sp[lo:hi, lo:hi] if contiguous else F.Sp[loc][:, loc]
    h = hashlib.blake2b(digest_size=16)
    for arr in (hrow, S0.indptr, S0.indices, S0.data, L0.indptr, L0.indices, Lh != 0):
        h.update(np.ascontiguousarray(arr).view(np.uint8).data)
    if hrows.size:
        Sh = F.Sp[hrows]
        Sh = Sh[:, lo:hi] if contiguous else Sh[:, loc]
        for arr in (Sh.indptr, Sh.indices, Sh.data):
            h.update(np.ascontiguousarray(arr).view(np.uint8).data)
    h.update(np.array(Lh.shape, dtype=np.int64).tobytes())
    return hrow, L0, Lh, h.digest()


def _subdomain_levels(hrow, L0, Lh, stage_cap, groups=None):
    """
    Sweep program of one subdomain as a list of levels (each a list of units of 8-row tiles); the arguments come
    from _subdomain_inputs.

    The local tree levels are collected into a few GROUPS of consecutive heights. Inside a group the
    triangular solve is replaced by the explicit inverse of the group's diagonal part (a block-diagonal matrix:
    one dense triangular block per subtree that lies inside the group), so a sweep needs two wide levels per
    group instead of two per tree level:
        forward    t_R = b_R - L[R, lower groups] z            (sparse rows, in place)
                   z_R = inv(L[R, R]) t_R                      (dense triangular blocks)
        backward   t_R = z_R - L[higher groups + halo, R]^T J  (sparse rows, in place)
                   J_R = inv(L[R, R])^T t_R
    A block of at most CHAIN_ROWS rows is applied in place by one warp (its row tiles in dependency order);
    larger blocks go through the staging rows, in passes of at most stage_cap rows ordered so that no pass
    reads a row an earlier pass has replaced.

    The local rows are renumbered: by group, then by block of the group's inverse, so every dense block is a
    contiguous row range. Returns (levels, n_bwd, order, group_bounds); order[i] = index into loc of local row i.
    """
    n = hrow.size
    if n == 0:
        return [], 0, np.zeros(0, dtype=np.int64), []
    n_halo = Lh.shape[0]
    H = int(hrow.max())
    if groups is None:
        groups = _auto_groups(L0, hrow, stage_cap)
    cuts = sorted(set(int(g) for g in groups if 0 <= int(g) < H))
    gid = np.searchsorted(np.asarray(cuts, dtype=np.int64), hrow, side="left") if cuts else np.zeros(n, dtype=np.int64)
    n_groups = len(cuts) + 1
    # blocks of each group's inverse = connected components of L restricted to the group
    comp = np.zeros(n, dtype=np.int64)
    for g in range(n_groups):
        R = np.flatnonzero(gid == g)
        if R.size == 0:
            continue
        sub = L0[R][:, R]
        _, lab = scipy.sparse.csgraph.connected_components(sub + sub.T, directed=False)
        first = np.full(lab.max() + 1, n, dtype=np.int64)
        np.minimum.at(first, lab, R)
        comp[R] = first[lab]
    order = np.lexsort((np.arange(n), comp, gid))
    Lp = L0.toarray()[np.ix_(order, order)]                  # dense from here on: a few hundred rows
    assert not np.any(np.triu(Lp, k=1))
    gid_p, comp_p = gid[order], comp[order]
    Lh = Lh[:, order]
    ident = np.arange(n + n_halo, dtype=np.int64)
    gstart = np.searchsorted(gid_p, np.arange(n_groups + 1))
    fwd, bwd = [], []
    for g in range(n_groups):
        a, b = int(gstart[g]), int(gstart[g + 1])
        if a == b:
            continue
        Linv = _tri_inverse(Lp[a:b, a:b])
        cstart = np.flatnonzero(np.concatenate(([True], comp_p[a + 1:b] != comp_p[a:b - 1], [True])))
        chains_f, chains_b, staged_f, staged_b = [], [], [], []
        for ci in range(cstart.size - 1):
            c0, c1 = int(cstart[ci]), int(cstart[ci + 1])            # relative to a
            starts = list(range(c0, c1, 8))
            tf, tb = [], []
            for t0 in starts:
                t1 = min(c1, t0 + 8)
                Vf = Linv[t0:t1, c0:t1]
                cf = np.flatnonzero(np.any(Vf != 0, axis=0))
                tf.append(dict(row0=a + t0, V=Vf[:, cf], cols=ident[a + c0 + cf], flags=0))
                Vb = Linv[t0:c1, t0:t1].T
                cb = np.flatnonzero(np.any(Vb != 0, axis=0))
                tb.append(dict(row0=a + t0, V=Vb[:, cb], cols=ident[a + t0 + cb], flags=0))
            if c1 - c0 <= CHAIN_ROWS:
                chains_f.append(tf[::-1])       # lower triangular: a row group reads the rows above it -> last group first
                chains_b.append(tb)
            else:
                staged_f.extend(tf)
                staged_b.extend(tb)

        def passes(tiles, descending):
            tiles = sorted(tiles, key=lambda t: -t["row0"] if descending else t["row0"])
            out, cur, rows = [], [], 0
            for t in tiles:
                t["flags"] = TILE_STAGED
                nr = t["V"].shape[0]
                if cur and rows + nr > stage_cap:
                    out.append(cur)
                    cur, rows = [], 0
                cur.append([t])
                rows += nr
            if cur:
                out.append(cur)
            return out
        pf, pb = passes(staged_f, True), passes(staged_b, False)
        inv_f = [chains_f + (pf[0] if pf else [])] + pf[1:]
        inv_b = [chains_b + (pb[0] if pb else [])] + pb[1:]
        fwd.append((_tiles_dense(Lp[a:b, :a], a, ident) if a > 0 else [], inv_f))
        above = np.concatenate((Lp[b:, a:b].T, Lh[:, a:b].T), axis=1) if (b < n or n_halo) else None
        bwd.append((_tiles_dense(above, a, ident[b:]) if above is not None else [], inv_b))
    levels = []
    for (ta, inv) in reversed(bwd):
        levels.append(ta)
        levels.extend(inv)
    levels = [u for u in levels if u]
    n_bwd = len(levels)
    for (ta, inv) in fwd:
        levels.append(ta)
        levels.extend(inv)
    if n_halo:
        levels.append(_tiles_dense(Lh, n, ident))
    levels = levels[:n_bwd] + [u for u in levels[n_bwd:] if u]
    return levels, n_bwd, order, cuts


def _auto_groups(L0, hrow, stage_cap):
    """Group boundaries: the lowest group takes as many tree levels as keep its subtrees within CHAIN_ROWS rows
    (applied in place by single warps); everything above forms one group (staged, usually a single pass)."""
    H = int(hrow.max())
    best = -1
    for hb in range(H):
        R = np.flatnonzero(hrow <= hb)
        sub = L0[R][:, R]
        _, lab = scipy.sparse.csgraph.connected_components(sub + sub.T, directed=False)
        if np.bincount(lab).max() <= CHAIN_ROWS:
            best = hb
        else:
            break
    return [best] if best >= 0 else []


UP_PLANE_SHIFT = 28                     # row code of the upper program: plane << 28 | row
PLANE_R, PLANE_Z, PLANE_J, PLANE_S = 0, 1, 2, 3     # right-hand side / forward result / solution of the separator rows; scratch
UP_SPLIT_MIN_COLS = 128                 # fewest columns of a column group of a split product
UP_TILES = 16                           # most 8-row tiles of a task of the upper program (one warp each)
UP_MERGE_ROWS = 96                      # separator blocks up to this size get one merged phase per depth and sweep
UP_WARP_COST = 512                      # tasks of at most this many A fragments (and 4 tiles) are run by single warps
UP_BLOCK_KSTEPS = 64                    # k-steps of a block task are a multiple of this (4 ring turns x up to 16 K slots)
UP_KSTEPS = 16                          # k-steps of a task are padded to a multiple of this (any K split 16 / 2^i works)


def _up_code(plane, rows):
    return (np.asarray(rows, dtype=np.int64) | (plane << UP_PLANE_SHIFT)).astype(np.int32)


class _UpperBuilder:
    """Collects the tasks of the upper program phase by phase (see JJSubdomainPlan in include/jjstep.h).
    A task computes  out[rows] = V . X[cols]  for at most 8 * RB consecutive rows of one plane; its A fragments are
    stored per 8-row tile as [k-step][lane = row * 4 + kk] = V[row, 4 k + kk]."""

    def __init__(self, pad_code, scratch_rows=0):
        self.RB, self.KB, self.pad_code = UP_TILES, UP_KSTEPS, pad_code
        self.scratch_rows, self.scratch_next, self.pending = scratch_rows, 0, []
        self.phases = []                 # list of lists of (cost, out code, rows, nk, cols, A)
        self.cur = None
        self.nnz = self.stored = 0       # factor entries applied / values stored (padding, unions of column sets)

    def begin_phase(self):
        self.cur = []
        self.scratch_next, self.pending = 0, []

    def end_phase(self):
        if self.cur:
            self.phases.append(self.cur)
        if self.pending:
            # the column groups of the split products of this phase are summed in a phase of their own
            self.cur = []
            for args in self.pending:
                self.task(*args)
            self.phases.append(self.cur)
        self.cur, self.pending = None, []

    def product(self, out_plane, row0, V, col_codes, self_plane=None, n_split=1):
        """out[rows] = V . X[cols] (+ self[rows]). n_split > 1 cuts the columns into that many groups whose partial
        products go to the scratch plane, one (task, chunk) pair each, and are summed - in a fixed order - by 8-row
        tasks of the phase that end_phase appends: parallelism for a few large separator blocks times a few problem
        chunks without shrinking the row blocks (which would gather every operand row for two row tiles only)."""
        nr, K = V.shape
        n_split = max(1, min(int(n_split), K // UP_SPLIT_MIN_COLS))
        if n_split > 1 and self.scratch_next + n_split * nr > self.scratch_rows:
            n_split = 1
        if n_split == 1:
            if self_plane is not None:
                V = np.concatenate((V, np.eye(nr)), axis=1)
                col_codes = np.concatenate((col_codes, _up_code(self_plane, np.arange(row0, row0 + nr))))
            self.task(out_plane, row0, V, col_codes)
            return
        edges = (np.linspace(0, K, n_split + 1) // 4 * 4).astype(int)
        edges[-1] = K
        srow = []
        for c in range(n_split):
            srow.append(self.scratch_next)
            self.task(PLANE_S, self.scratch_next, V[:, edges[c]:edges[c + 1]], col_codes[edges[c]:edges[c + 1]])
            self.scratch_next += nr
        for i0 in range(0, nr, 8):
            m = min(8, nr - i0)
            codes = [_up_code(PLANE_S, np.arange(s0 + i0, s0 + i0 + m)) for s0 in srow]
            if self_plane is not None:
                codes.append(_up_code(self_plane, np.arange(row0 + i0, row0 + i0 + m)))
            self.pending.append((out_plane, row0 + i0, np.tile(np.eye(m), (1, len(codes))), np.concatenate(codes)))

    def task(self, out_plane, row0, V, col_codes):
        nr, K = V.shape
        assert 1 <= nr <= 8 * self.RB and K == len(col_codes) and K > 0
        tiles = -(-nr // 8)
        nk = -(-K // (4 * self.KB)) * self.KB
        warp = tiles * nk <= UP_WARP_COST and nr <= 32
        if not warp:
            nk = -(-K // (4 * UP_BLOCK_KSTEPS)) * UP_BLOCK_KSTEPS      # a block task: whole turns of the A ring per warp
        self.nnz += int(np.count_nonzero(V))
        self.stored += tiles * 8 * nk * 4
        A = np.zeros((tiles * 8, nk * 4))
        A[:nr, :K] = V
        A = np.ascontiguousarray(A.reshape(tiles, 8, nk, 4).transpose(0, 2, 1, 3)).ravel()
        cols = np.full(nk * 4, self.pad_code, dtype=np.int32)
        cols[:K] = col_codes
        self.cur.append((tiles * nk, int(_up_code(out_plane, row0)), nr, nk, cols, A, warp))

    def finish(self, n_fwd):
        """-> dict of flat arrays; tasks of a phase sorted by decreasing cost (the device deals them out round-robin)."""
        phase_ptr, phase_split, hdr, aoff, cols, vals = [0], [], [], [], [], []
        co = vo = 0
        for ph in self.phases:
            # block tasks (panels in shared memory, 16 warps) first, then the tasks small enough for a single warp
            ordered = sorted((t for t in ph if not t[6]), key=lambda t: -t[0])
            phase_split.append(len(hdr) + len(ordered))
            ordered += sorted((t for t in ph if t[6]), key=lambda t: -t[0])
            for (cost, out, nr, nk, c, A, _) in ordered:
                hdr.append((out, nr, nk, co))
                aoff.append(vo)
                cols.append(c)
                vals.append(A)
                co += c.size
                vo += A.size
            phase_ptr.append(len(hdr))
        assert co < 2 ** 31
        return dict(RB=self.RB, KB=self.KB, n_fwd=n_fwd, n_bwd=len(self.phases) - n_fwd, nnz=self.nnz, stored=self.stored,
                    phase_ptr=np.asarray(phase_ptr, dtype=np.int32),
                    phase_split=np.asarray(phase_split + [0], dtype=np.int32),
                    task=np.asarray(hdr, dtype=np.int32).reshape(-1, 4),
                    task_aoff=np.asarray(aoff, dtype=np.int64),
                    cols=np.concatenate(cols) if cols else np.zeros(0, dtype=np.int32),
                    A=np.concatenate(vals) if vals else np.zeros(0))


def _phase_rows(blocks, n_chunks, n_sm=148):
    """Rows per task for the separator blocks of one tree depth: as many as 16 warps can take (128), fewer when that
    would leave SMs without a (task, chunk) pair - near the root there are few, large blocks."""
    for rb in (16, 8, 4):
        if sum(-(-(b1 - b0) // (8 * rb)) for (b0, b1) in blocks) * n_chunks >= 2 * n_sm:
            return 8 * rb
    return 16


def _upper_program(F, top_rows, tt0, blk_of, n_chunks, n_up_pad):
    """
    Sweep program of the UPPER separators (top rows [0, tt0) in the top numbering) and the elimination of the dense
    top of the top (rows [tt0, n_top)) from them. Separator blocks of one tree depth are mutually independent, so a
    sweep takes two phases per depth, every phase a list of independent gathered dense products:

      forward, deepest first    a)  t_B = r_B - L[B, below] z          (r plane, in place)
                                b)  z_B = inv(L[B, B]) t_B             (r -> z plane)
      then                          r_tt -= L[tt, upper] z             (r plane, in place; the dense inverse follows)
      backward, shallowest first a) t_B = z_B - L[above, B]^T J        (z plane, in place)
                                b)  J_B = inv(L[B, B])^T t_B           (z -> J plane)

    Depths whose blocks are all small are done in ONE phase per sweep (a and b multiplied out on the host), and
    small tasks are marked for single warps: a phase lists its block tasks first, then its warp tasks.
    """
    n_top = top_rows.size
    ub = _UpperBuilder(int(_up_code(PLANE_R, n_up_pad - 1)), scratch_rows=n_up_pad)
    if tt0 == 0:
        return ub.finish(0)
    Lt = F.Lc[top_rows][:, top_rows].tocsr()
    Lt.sort_indices()
    assert scipy.sparse.triu(Lt, k=1).nnz == 0
    LtT = Lt.T.tocsr()
    LtT.sort_indices()
    bu = blk_of[top_rows[:tt0]]
    starts = np.flatnonzero(np.concatenate(([True], bu[1:] != bu[:-1])))
    ends = np.concatenate((starts[1:], [tt0]))
    bdepth = F.depth[bu[starts]].astype(np.int64)
    depths = np.unique(bdepth)
    dinv = {}

    def block_inv(b0, b1):
        if b0 not in dinv:
            dinv[b0] = _tri_inverse(Lt[b0:b1, b0:b1].toarray())
        return dinv[b0]

    def groups_of(blocks, splittable=False):
        """row groups (r0, r1, b0, b1) of the tasks of a phase and the number of column groups of each product"""
        RT = _phase_rows(blocks, n_chunks)
        n_split = 1
        if splittable and RT < 64:
            # a few large blocks and a few problem chunks: keep 128-row tasks and split their columns instead
            RT = 128
            n_groups = sum(-(-(b1 - b0) // RT) for (b0, b1) in blocks)
            n_split = -(-2 * 148 // (n_groups * n_chunks))
        return [(r0, min(b1, r0 + RT), b0, b1) for (b0, b1) in blocks for r0 in range(b0, b1, RT)], n_split

    def dense_cols(M, r0, r1, lo, hi):
        """columns in [lo, hi) that rows r0:r1 of the CSR matrix M touch, and the dense block over them (straight from
        the CSR arrays: thousands of these per plan, scipy's fancy indexing costs more than the work)"""
        ptr = M.indptr[r0:r1 + 1]
        idx, val = M.indices[ptr[0]:ptr[-1]], M.data[ptr[0]:ptr[-1]]
        keep = np.flatnonzero((idx >= lo) & (idx < hi))
        if keep.size == 0:
            return np.zeros(0, dtype=idx.dtype), np.zeros((r1 - r0, 0))
        kept = idx[keep]
        lo = int(kept.min())                         # (the touched columns usually span a small part of the window)
        rel = kept - lo
        mark = np.zeros(int(kept.max()) - lo + 1, dtype=bool)     # a bitmap of the column span instead of a sort of the entries
        mark[rel] = True
        cols = (np.flatnonzero(mark) + lo).astype(idx.dtype)
        pos = (np.cumsum(mark) - 1)[rel]
        rows = np.repeat(np.arange(r1 - r0), np.diff(ptr))[keep]
        D = np.zeros((r1 - r0, cols.size))
        D[rows, pos] = val[keep]                     # (canonical CSR: one entry per position)
        return cols, D

    def merged(blocks):
        """small separators: one phase per depth, z_B = inv(L_BB) r_B - (inv(L_BB) L[B, below]) z (the product is as
        dense as L[B, below] itself when the block is a few tiles), instead of two phases with a barrier between"""
        return max(b1 - b0 for (b0, b1) in blocks) <= UP_MERGE_ROWS

    # ---- forward
    for dd in depths[::-1]:
        blocks = [(int(starts[i]), int(ends[i])) for i in np.flatnonzero(bdepth == dd)]
        if merged(blocks):
            grp, _ = groups_of(blocks)
            ub.begin_phase()
            for (r0, r1, b0, b1) in grp:
                Di = block_inv(b0, b1)[r0 - b0:r1 - b0]
                cols, Lb = dense_cols(Lt, b0, b1, 0, b0)            # all rows of the block: Di couples them
                V = np.concatenate((Di[:, :r1 - b0], -(Di @ Lb)), axis=1)
                ub.task(PLANE_Z, r0, V, np.concatenate((_up_code(PLANE_R, np.arange(b0, r1)), _up_code(PLANE_Z, cols))))
            ub.end_phase()
            continue
        grp, n_split = groups_of(blocks, splittable=True)
        ub.begin_phase()
        for (r0, r1, b0, b1) in grp:
            cols, Lb = dense_cols(Lt, r0, r1, 0, b0)
            if cols.size:
                ub.product(PLANE_R, r0, -Lb, _up_code(PLANE_Z, cols), self_plane=PLANE_R, n_split=n_split)
        ub.end_phase()
        ub.begin_phase()
        for (r0, r1, b0, b1) in grp:
            Di = block_inv(b0, b1)
            ub.product(PLANE_Z, r0, Di[r0 - b0:r1 - b0, :r1 - b0], _up_code(PLANE_R, np.arange(b0, r1)), n_split=n_split)
        ub.end_phase()
    if n_top > tt0:
        grp, n_split = groups_of([(tt0, n_top)], splittable=True)
        ub.begin_phase()
        for (r0, r1, _b0, _b1) in grp:
            cols, Lb = dense_cols(Lt, r0, r1, 0, tt0)
            if cols.size:
                ub.product(PLANE_R, r0, -Lb, _up_code(PLANE_Z, cols), self_plane=PLANE_R, n_split=n_split)
        ub.end_phase()
    n_fwd = len(ub.phases)
    # ---- backward
    for dd in depths:
        blocks = [(int(starts[i]), int(ends[i])) for i in np.flatnonzero(bdepth == dd)]
        if merged(blocks):
            grp, _ = groups_of(blocks)
            ub.begin_phase()
            for (r0, r1, b0, b1) in grp:
                DiT = block_inv(b0, b1).T[r0 - b0:r1 - b0]          # rows r0:r1 of inv(L_BB)^T, columns = block rows
                cols, Lb = dense_cols(LtT, b0, b1, b1, n_top)
                V = np.concatenate((DiT[:, r0 - b0:], -(DiT @ Lb)), axis=1)
                ub.task(PLANE_J, r0, V, np.concatenate((_up_code(PLANE_Z, np.arange(r0, b1)), _up_code(PLANE_J, cols))))
            ub.end_phase()
            continue
        grp, n_split = groups_of(blocks, splittable=True)
        ub.begin_phase()
        for (r0, r1, b0, b1) in grp:
            cols, Lb = dense_cols(LtT, r0, r1, b1, n_top)
            if cols.size:
                ub.product(PLANE_Z, r0, -Lb, _up_code(PLANE_J, cols), self_plane=PLANE_Z, n_split=n_split)
        ub.end_phase()
        ub.begin_phase()
        for (r0, r1, b0, b1) in grp:
            Di = block_inv(b0, b1)
            ub.product(PLANE_J, r0, Di.T[r0 - b0:r1 - b0, r0 - b0:], _up_code(PLANE_Z, np.arange(r0, b1)), n_split=n_split)
        ub.end_phase()
    return ub.finish(n_fwd)


def subdomain_plan(F, junc_face, d, NG, n_warps=RES_WARPS, groups=None, tt_max=None, n_chunks=1):
    """
    F         : factor.Factor of the permuted cycle-space system
    junc_face : (Nj, 2) permuted faces of every junction, -1 none (CircuitTables.junc_face)
    d         : cut depth, P = 2^d subdomains; None: the subtrees of an ordering made with n_parts (F.blk_part)
    NG        : problem groups of 8 per chunk (PC = 8 * NG problems share one pass over the factor)
    groups    : tree heights after which a new level group starts (see _subdomain_levels); None = automatic
    tt_max    : the tree levels nearest the root whose rows number at most this many are solved by ONE dense inverse
                (the top of the top); the separators between them and the subdomains get the upper program
    n_chunks  : problem chunks the plan will run with (sizes the tasks of the upper program: a phase wants at least
                two (task, chunk) pairs per SM)
    """
    assert NG in (1, 2, 4, 8)
    n, nb = F.n, F.nb
    PC = 8 * NG
    if tt_max is None:
        # measured on B200: 1024 for wide batches (cfg3 / cfg4 at 512 problems lose 3 % at 2048 and 25 % at 4096: the dense
        # product costs n_tt^2 per problem), 4096 for narrow ones (cfg5 at 64 problems gains 4 %: the inverse streams from
        # HBM once per step whatever the width, and replaces the most serial upper phases)
        tt_max = int(os.environ.get("JJ_TT_MAX", "4096" if n_chunks * 8 * NG <= 128 else "1024"))
    sizes = np.diff(F.bptr)
    blk_of = np.repeat(np.arange(nb), sizes)
    if d is None:
        # the ordering was made with n_parts equal subtrees: they are the subdomains
        assert F.blk_part is not None
        blk_sub = np.asarray(F.blk_part, dtype=np.int64)
        P = int(blk_sub.max()) + 1
    else:
        P = 1 << d
        top_blk = F.depth < d
        blk_sub = np.where(top_blk, -1, F.dom >> np.maximum(F.depth - d, 0)).astype(np.int64)
    row_sub = blk_sub[blk_of]
    # the top of the top: all separator blocks of depth < d_tt (an ancestor-closed set: it can be eliminated last)
    tb = np.flatnonzero(blk_sub < 0)
    d_tt = 0
    if tb.size:
        rows_at = np.bincount(F.depth[tb], weights=sizes[tb], minlength=int(F.depth[tb].max()) + 1)
        d_tt = int(np.searchsorted(np.cumsum(rows_at), tt_max, side="right"))
    row_tt = ((blk_sub < 0) & (F.depth < d_tt))[blk_of]
    su_rows = np.flatnonzero((row_sub < 0) & ~row_tt)
    tt_rows = np.flatnonzero(row_tt)
    top_rows = np.concatenate((su_rows, tt_rows))
    n_top, tt0, n_tt = int(top_rows.size), int(su_rows.size), int(tt_rows.size)
    tix = np.full(n, -1, dtype=np.int64)
    tix[top_rows] = np.arange(n_top)

    # (subdomain, top row) couplings through the factor: entries of the top rows of Lc in local columns
    # (local rows only ever couple to their own subdomain: every subdomain is a subtree of the elimination tree)
    Ltr = F.Lc[top_rows]
    tr_row = np.repeat(np.arange(n_top), np.diff(Ltr.indptr))
    rs_c = row_sub[Ltr.indices]
    m = rs_c >= 0
    nt1 = max(n_top, 1)
    lkeys = np.unique(rs_c[m] * nt1 + tr_row[m])
    del Ltr, tr_row, rs_c, m
    # subdomains coupled to each top row (CSR), for the junctions that lie between two top faces
    lk_s, lk_k = lkeys // nt1, lkeys % nt1
    o = np.argsort(lk_k, kind="stable")
    sot_ptr = np.searchsorted(lk_k[o], np.arange(n_top + 1))
    sot = lk_s[o]

    # junction ownership: the subdomain of its local face; junctions between top faces go to a coupled
    # subdomain with the fewest junctions so far
    jf = np.asarray(junc_face, dtype=np.int64)
    Nj = jf.shape[0]
    has = jf >= 0
    fs = np.where(has, row_sub[np.maximum(jf, 0)], -1)
    owner = np.maximum(fs[:, 0], fs[:, 1])
    both = (fs[:, 0] >= 0) & (fs[:, 1] >= 0)
    assert np.all(fs[both, 0] == fs[both, 1]), "faces sharing a junction ended up in different subdomains"
    load = np.bincount(owner[owner >= 0], minlength=P).astype(np.int64)
    for j in np.flatnonzero(owner < 0):
        cand = set()
        for k in range(2):
            if has[j, k]:
                t = tix[jf[j, k]]
                cand.update(sot[sot_ptr[t]:sot_ptr[t + 1]].tolist())
        cand = sorted(cand) if cand else list(range(P))
        s = min(cand, key=lambda q: (load[q], q))
        owner[j] = s
        load[s] += 1
    # a subdomain also carries the top faces of the junctions it owns (partial face sums, back-projection)
    jt = has & (np.where(has, row_sub[np.maximum(jf, 0)], 0) < 0)
    jkeys = (np.repeat(owner[:, None], 2, axis=1)[jt]) * nt1 + tix[jf[jt]]
    keys = np.unique(np.concatenate((lkeys, jkeys)))            # sorted by (subdomain, top row)
    key_s, key_k = keys // nt1, keys % nt1

    plan = SubdomainPlan()
    plan.P, plan.NG, plan.PC, plan.d = P, NG, PC, d
    plan.n_top, plan.top_rows = n_top, top_rows.astype(np.int32)
    plan.tt0, plan.n_tt = tt0, n_tt
    plan.n_tt_pad = (n_tt + 31) // 32 * 32
    plan.n_up_pad = (max(n_top + 1, tt0 + plan.n_tt_pad) + 31) // 32 * 32
    if plan.n_tt_pad > 8192:
        raise ValueError("subdomain plan: %d rows are too many for the dense top product" % n_tt)
    if plan.n_up_pad >= (1 << UP_PLANE_SHIFT):
        raise ValueError("subdomain plan: %d separator rows exceed the row codes of the upper program" % n_top)
    order_l = np.argsort(row_sub, kind="stable")
    lptr = np.searchsorted(row_sub[order_l], np.arange(P + 1))
    loc = [order_l[lptr[s]:lptr[s + 1]] for s in range(P)]
    plan.hptr = np.searchsorted(key_s, np.arange(P + 1)).astype(np.int32)
    halo = [key_k[plan.hptr[s]:plan.hptr[s + 1]] for s in range(P)]
    plan.n_loc = np.diff(lptr).astype(np.int32)
    plan.n_halo = np.diff(plan.hptr).astype(np.int32)
    plan.n_loc_max = int(plan.n_loc.max())
    plan.n_rows = (int((plan.n_loc + plan.n_halo).max()) + 7) // 8 * 8 + 8
    if plan.n_rows * PC > 65536:
        raise ValueError("subdomain plan: %d rows x %d problems exceed the 16-bit element codes" % (plan.n_rows, PC))
    plan.n_slots = int(plan.hptr[-1])
    plan.halo_top = key_k.astype(np.int32)
    plan.row_sub = row_sub

    def halo_row(s, k):
        """shared-memory row of top row k in subdomain s (vectorised)."""
        pos = np.searchsorted(keys, s * nt1 + k)
        assert np.all(keys[np.minimum(pos, keys.size - 1)] == s * nt1 + k)
        return plan.n_loc[s] + (pos - plan.hptr[s])
    plan.halo_row = halo_row

    # ---- per-subdomain sweep programs: backward levels first (they open a time step), then forward
    plan.prog, plan.n_bwd = [], []
    plan.group_bounds = []
    # room for headers, cursors and the amplitude cache; half blocks (8 warps, two to an SM) have 112 KB each
    smem_limit = (227 * 1024 - 20 * 1024) if n_warps >= 16 else (112 * 1024 - 8 * 1024)
    stage_cap = (smem_limit // 8 - plan.n_rows * PC) // (PC + 2)
    if stage_cap < 8:
        raise ValueError("subdomain plan: the right-hand sides leave no room for the staging rows")
    stage_cap = min(stage_cap, 4096)
    vrow_loc = np.full(n, -1, dtype=np.int64)            # shared-memory row of every local face in its subdomain
    cache = {}                                           # digest of a subdomain's inputs -> its program
    # the program of the upper separators depends on the factor and the cut only: it is built on a second thread while
    # this one goes through the subdomains (numpy / scipy / LAPACK release the interpreter lock in their kernels)
    upper_job = _background(_upper_program, F, top_rows, tt0, blk_of, max(1, int(n_chunks)), plan.n_up_pad)
    for s in range(P):
        if loc[s].size == 0:             # a cut deeper than a small circuit's tree leaves subtrees without rows
            hit = cache.setdefault(b"empty", (_pack_levels([], NG, n_warps, 0), 0, np.zeros(0, dtype=np.int64), [],
                                              np.zeros(0), np.zeros((0, 0))))
            plan.prog.append(hit[0])
            plan.n_bwd.append(0)
            plan.group_bounds.append([])
            continue
        hrow, L0, Lh, digest = _subdomain_inputs(F, loc[s], top_rows[halo[s]], blk_of)
        hit = cache.get(digest)
        if hit is not None:
            # congruent to an earlier subdomain: its factor blocks must agree to rounding (guards the digest)
            scale = 1e-12 * float(np.max(np.abs(L0.data)))
            if np.max(np.abs(hit[4] - L0.data)) > scale or (Lh.size and np.max(np.abs(hit[5] - Lh)) > scale):
                hit = None
        if hit is None:
            levels, n_bwd, order, gb = _subdomain_levels(hrow, L0, Lh, stage_cap, groups)
            hit = cache[digest] = (_pack_levels(levels, NG, n_warps, n_bwd), n_bwd, order, gb, L0.data, Lh)
        prog, n_bwd, order, gb = hit[:4]
        vrow_loc[loc[s][order]] = np.arange(loc[s].size)
        plan.prog.append(prog)
        plan.n_bwd.append(n_bwd)
        plan.group_bounds.append(gb)
    plan.n_distinct_programs = len(cache)
    plan.vrow_loc = vrow_loc
    plan.n_bwd = np.asarray(plan.n_bwd, dtype=np.int32)
    plan.stage_rows = max(p["stage_rows"] for p in plan.prog)
    if n_top:
        # the top product and the upper phases keep a ring of 4 stages x 8 KB (or 4 KB, 2 KB) in the staging rows
        for ring in (32768, 16384, 8192):
            need = -(-ring // ((PC + 2) * 8))
            if need <= stage_cap:
                plan.stage_rows = max(plan.stage_rows, need)
                break

    # ---- top of the top: explicit inverse of its Schur complement, packed as FP64 MMA A fragments
    nTp = plan.n_tt_pad
    if n_tt:
        LTT = F.Lc[tt_rows][:, tt_rows].toarray()
        Linv = _tri_inverse(LTT)
        Sinv = Linv.T @ Linv
        SP = np.zeros((nTp, nTp))
        SP[:n_tt, :n_tt] = Sinv
        plan.Sinv = Sinv
        plan.Sinv_packed = np.ascontiguousarray(SP.reshape(nTp // 8, 8, nTp // 4, 4).transpose(0, 2, 1, 3)).ravel()
    else:
        plan.Sinv = np.zeros((0, 0))
        plan.Sinv_packed = np.zeros(0)
    # ---- upper separators between the subdomains and the top of the top
    plan.upper = upper_job()
    if plan.upper["n_fwd"] + plan.upper["n_bwd"] > 0:
        # the upper phases gather into two panel buffers of at least 256 rows and reduce over 16 warps in shared memory
        need = max(2 * (128 if NG >= 8 else 256) * PC, 16 * NG * 64) - plan.n_rows * PC
        plan.stage_rows = max(plan.stage_rows, -(-need // (PC + 2)))
    # assembly of r_top: slots (subdomain halo rows) of every top row
    order = np.argsort(plan.halo_top, kind="stable")
    plan.tslot = order.astype(np.int32)
    plan.tptr = np.searchsorted(plan.halo_top[order], np.arange(n_top + 1)).astype(np.int32)

    # ---- junction / face tables in device junction order (grouped by owner)
    plan.owner = owner
    jorder = np.lexsort((np.arange(Nj), owner))
    plan.junc_orig = jorder.astype(np.int32)
    jdev = np.empty(Nj, dtype=np.int64)
    jdev[jorder] = np.arange(Nj)
    plan.jdev = jdev
    plan.junc_ptr = np.searchsorted(owner[jorder], np.arange(P + 1)).astype(np.int32)
    plan.junc_row = np.ascontiguousarray(plan.rows_of(np.repeat(owner[:, None], 2, axis=1), jf)[jorder].astype(np.int32))
    return plan


def _rows_of(plan, s, g):
    """Shared-memory row of permuted face g in subdomain s (arrays of equal shape; g = -1 gives -1)."""
    s, g = np.asarray(s, dtype=np.int64), np.asarray(g, dtype=np.int64)
    out = np.full(g.shape, -1, dtype=np.int64)
    ok = g >= 0
    gs = plan.row_sub[np.maximum(g, 0)]
    is_loc = ok & (gs >= 0)
    assert np.all(gs[is_loc] == s[is_loc]), "a face is used by a subdomain that does not own it"
    out[is_loc] = plan.vrow_loc[g[is_loc]]
    is_top = ok & (gs < 0)
    if is_top.any():
        tix = np.full(plan.row_sub.size, -1, dtype=np.int64)
        tix[plan.top_rows] = np.arange(plan.n_top)
        out[is_top] = plan.halo_row(s[is_top], tix[g[is_top]])
    return out


SubdomainPlan.rows_of = _rows_of


def face_tables(plan, face_ptr, face_junc, face_sign, junc_sign, c0):
    """Fixed-width per-(subdomain, row) lists of (device junction, sign / c0) for the face projection
    b = A (x / c0 - theta_s) (reference: time_evolution.py:560-569); a halo row lists only the junctions its
    subdomain owns, the partial sums meet in the top assembly."""
    P, n_rows = plan.P, plan.n_rows
    Nf = len(face_ptr) - 1
    g_of = np.repeat(np.arange(Nf), np.diff(face_ptr))
    j_of = np.asarray(face_junc, dtype=np.int64)
    r_of = plan.owner[j_of]
    row_of = plan.rows_of(r_of, g_of)
    assert np.all(row_of >= 0)
    key = np.lexsort((j_of, row_of, r_of))
    flat = (r_of * n_rows + row_of)[key]
    counts = np.bincount(flat, minlength=P * n_rows)
    K = max(4, (int(counts.max()) + 3) // 4 * 4) if counts.size else 4
    start = np.concatenate(([0], np.cumsum(counts)))[:-1]
    pos = np.arange(flat.size) - start[flat]
    ell_j = np.full((P * n_rows, K), -1, dtype=np.int32)
    ell_c = np.zeros((P * n_rows, K))
    ell_j[flat, pos] = plan.jdev[j_of[key]]
    ell_c[flat, pos] = np.asarray(face_sign, dtype=np.double)[key] / c0[j_of[key]]
    fidx = np.full((P, n_rows), -1, dtype=np.int32)
    g = np.flatnonzero(plan.row_sub >= 0)
    fidx[plan.row_sub[g], plan.vrow_loc[g]] = g
    plan.face_K = K
    plan.face_ell_j = ell_j.reshape(P, n_rows, K)
    plan.face_ell_c = ell_c.reshape(P, n_rows, K)
    plan.face_fidx = fidx
    plan.junc_sign = np.ascontiguousarray(np.asarray(junc_sign)[plan.junc_orig].astype(np.int8))
    return plan


# ----------------------------------------------------------------------------------------------
# host interpreter (CPU tests of the plan; mirrors the device data flow)
# ----------------------------------------------------------------------------------------------
def _run_level_host(ps, v, level, NG):
    nw = ps["n_warps"]
    rec = ps["stream"].reshape(-1, STEP_BYTES)
    PC = 8 * NG
    staged = []
    for w in range(nw):
        idx = level * nw + w
        s = ps["ws_ptr"][idx]
        for t in range(ps["wt_ptr"][idx, 0], ps["wt_ptr"][idx, 1]):
            h0, h1 = int(ps["thdr"][t, 0]), int(ps["thdr"][t, 1])
            row0, nr, fl = h0 & 0xffff, ((h0 >> 16) & 7) + 1, (h0 >> 19) & 3
            g0, ng = (h0 >> 21) & 15, ((h0 >> 25) & 15) + 1
            st = h1 & 0xffff
            vals = rec[s:s + st, :256].copy().view(np.float64).reshape(st, 8, 4)                      # [step][row][kk]
            codes = rec[s:s + st, 256:].copy().view(np.uint16).reshape(st, 8, 4).astype(np.int64)    # [step][n][kk]
            s += st
            cols = codes[:, 0, :] // PC                                                              # [step][kk]
            assert np.array_equal(elem_code(cols, NG).transpose(0, 2, 1), codes)
            pcols = slice(g0 * 8, (g0 + ng) * 8)
            acc = np.einsum("srk,skn->rn", vals, v[cols][:, :, pcols])[:nr]
            if fl & TILE_SELF:
                acc = acc + v[row0:row0 + nr, pcols]
            if fl & TILE_STAGED:
                staged.append((row0, nr, pcols, acc))
            else:
                v[row0:row0 + nr, pcols] = acc
    for (row0, nr, pcols, acc) in staged:
        v[row0:row0 + nr, pcols] = acc


def _run_upper_phase_host(up, U, ph):
    """One phase of the upper program on the host; U is (4, n_up_pad, PC). Tasks of a phase are independent: all
    products are formed before any row is written, as on the device (where a grid barrier ends the phase)."""
    res = []
    mask = (1 << UP_PLANE_SHIFT) - 1
    for t in range(up["phase_ptr"][ph], up["phase_ptr"][ph + 1]):
        out, nr, nk, co = (int(v) for v in up["task"][t])
        tiles = -(-nr // 8)
        ao = int(up["task_aoff"][t])
        A = up["A"][ao:ao + tiles * nk * 32].reshape(tiles, nk, 8, 4).transpose(0, 2, 1, 3).reshape(tiles * 8, nk * 4)
        cols = up["cols"][co:co + nk * 4].astype(np.int64)
        X = U[cols >> UP_PLANE_SHIFT, cols & mask]
        res.append((out >> UP_PLANE_SHIFT, out & mask, (A @ X)[:nr]))
    for (pl, r0, val) in res:
        U[pl, r0:r0 + val.shape[0]] = val


def apply_subdomain_plan_host(plan, b_perm):
    """Solve through the plan on the host. b_perm: (Nf, PC) right-hand side in PERMUTED numbering."""
    P, NG = plan.P, plan.NG
    assert b_perm.shape[1] == plan.PC
    vec = []
    ctop = np.zeros((plan.n_slots, plan.PC))
    row_of = np.where(plan.row_sub >= 0, plan.vrow_loc, -1)
    for s in range(P):
        v = np.zeros((plan.n_rows, plan.PC))
        loc = np.flatnonzero(plan.row_sub == s)
        v[row_of[loc]] = b_perm[loc]
        ps = plan.prog[s]
        for l in range(plan.n_bwd[s], ps["n_levels"]):
            _run_level_host(ps, v, l, NG)
        ctop[plan.hptr[s]:plan.hptr[s + 1]] = v[plan.n_loc[s]: plan.n_loc[s] + plan.n_halo[s]]
        vec.append(v)
    U = np.zeros((4, plan.n_up_pad, plan.PC))
    U[PLANE_R, :plan.n_top] = b_perm[plan.top_rows]
    for k in range(plan.n_top):
        for sl in plan.tslot[plan.tptr[k]:plan.tptr[k + 1]]:
            U[PLANE_R, k] += ctop[sl]
    up = plan.upper
    for ph in range(up["n_fwd"]):
        _run_upper_phase_host(up, U, ph)
    t0, t1 = plan.tt0, plan.tt0 + plan.n_tt
    U[PLANE_J, t0:t1] = plan.Sinv @ U[PLANE_R, t0:t1]
    for ph in range(up["n_fwd"], up["n_fwd"] + up["n_bwd"]):
        _run_upper_phase_host(up, U, ph)
    assert not np.any(U[PLANE_R, plan.n_top:]), "the zero rows of the r plane were written"
    jtop = U[PLANE_J, :plan.n_top]
    out = np.zeros_like(b_perm)
    out[plan.top_rows] = jtop
    for s in range(P):
        v = vec[s]
        v[plan.n_loc[s]: plan.n_loc[s] + plan.n_halo[s]] = jtop[plan.halo_top[plan.hptr[s]:plan.hptr[s + 1]]]
        ps = plan.prog[s]
        for l in range(plan.n_bwd[s]):
            _run_level_host(ps, v, l, NG)
        loc = np.flatnonzero(plan.row_sub == s)
        out[loc] = v[row_of[loc]]
    return out
