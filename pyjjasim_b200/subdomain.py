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


def _tile_record(V, cols, NG):
    """320-byte stream steps of one tile: A fragment (lane = row*4 + kk) and element codes (lane = n*4 + kk)."""
    nr, nc = V.shape
    st = ((nc + 3) // 4 + RING_STEPS - 1) // RING_STEPS * RING_STEPS      # whole ring blocks (zero steps at the end)
    Vp = np.zeros((8, st * 4))
    Vp[:nr, :nc] = V
    cp = np.full(st * 4, cols[-1], dtype=np.int64)
    cp[:nc] = cols
    vals = Vp.reshape(8, st, 4).transpose(1, 0, 2).reshape(st, 32)
    codes = elem_code(cp.reshape(st, 4), NG)                       # (st, kk, n)
    assert codes.max() < 65536
    codes = codes.transpose(0, 2, 1).reshape(st, 32).astype(np.uint16)
    rec = np.zeros((st, STEP_BYTES), dtype=np.uint8)
    rec[:, :256] = np.ascontiguousarray(vals).view(np.uint8).reshape(st, 256)
    rec[:, 256:] = np.ascontiguousarray(codes).view(np.uint8).reshape(st, 64)
    return rec


def _pack_levels(levels, NG, n_warps, n_bwd=0):
    """levels: list of unit lists; a unit is a list of tiles dict(row0, V, cols, flags) executed in order by
    one warp. Returns the per-(level, warp) tile ranges and streams. Units longer than a warp's fair share of a
    level (and units of levels with fewer units than warps) are split over problem groups: a warp then handles
    ng < NG groups of the unit; the copies re-read the unit's (small) stream but run concurrently.
    Tiles and stream steps are emitted warp-major inside each sweep (levels [0, n_bwd) and [n_bwd, ...)), so the
    stream of a warp is contiguous across the levels of a sweep and its prefetch ring never drains."""
    n_levels = len(levels)
    wt_ptr = np.zeros((n_levels * n_warps, 2), dtype=np.int32)
    ws_ptr = np.zeros(n_levels * n_warps, dtype=np.int32)
    hdr, chunks, lstaged = [], [], []
    n_steps = n_vals = 0
    assigned = []                      # per level: per warp list of (unit index, g0, ng)
    for li, units in enumerate(levels):
        staged = 0
        for u in units:
            for t in u:
                t["rec"] = _tile_record(t["V"], t["cols"], NG)
                t["stage_off"] = 0
                if t["flags"] & TILE_STAGED:
                    t["stage_off"] = staged
                    staged += t["V"].shape[0]
        lstaged.append(staged)
        ucost = [sum(t["rec"].shape[0] for t in u) for u in units]
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
                        st = t["rec"].shape[0]
                        assert 0 <= t["row0"] < 65536 and st < 65536 and t["stage_off"] < 32768
                        hdr.append((t["row0"] | ((nr - 1) << 16) | (t["flags"] << 19) | (g0 << 21) | ((ng - 1) << 25),
                                    st | (t["stage_off"] << 16)))
                        chunks.append(t["rec"])
                        n_steps += st
                        if g0 == 0:
                            n_vals += nr * nc
                wt_ptr[li * n_warps + w, 1] = len(hdr)
    stream = np.concatenate(chunks).ravel() if chunks else np.zeros(0, dtype=np.uint8)
    return dict(n_levels=n_levels, n_warps=n_warps, wt_ptr=wt_ptr, ws_ptr=ws_ptr,
                thdr=np.asarray(hdr, dtype=np.int32).reshape(-1, 2), stream=stream, n_steps=int(n_steps),
                lstaged=np.asarray(lstaged, dtype=np.int32), stage_rows=int(max(lstaged) if lstaged else 0),
                vals=int(n_vals))


def _tiles_sparse(M, row0, col_index, flags=TILE_SELF):
    """8-row tiles of out[row0 + i] (+)= -M[i, :] src: M is CSR, col_index maps its columns to vector rows."""
    units = []
    M = scipy.sparse.csr_matrix(M)
    for t0 in range(0, M.shape[0], 8):
        sub = M[t0:t0 + 8]
        cols = np.unique(sub.indices[sub.data != 0]) if sub.nnz else np.zeros(0, dtype=np.int64)
        if cols.size == 0:
            continue
        units.append([dict(row0=int(row0 + t0), V=-sub[:, cols].toarray(), cols=col_index[cols], flags=flags)])
    return units


def _subdomain_levels(F, loc, hrows, blk_of, stage_cap, groups=None):
    """
    Sweep program of one subdomain as a list of levels (each a list of units of 8-row tiles).

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
    n = loc.size
    if n == 0:
        return [], 0, np.zeros(0, dtype=np.int64), []
    hrow = F.height[blk_of[loc]].astype(np.int64)
    H = int(hrow.max())
    L0 = scipy.sparse.csr_matrix(F.Lc[loc][:, loc])
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
    Lp = L0[order][:, order].tocsr()
    assert scipy.sparse.triu(Lp, k=1).nnz == 0
    gid_p, comp_p = gid[order], comp[order]
    Lh = scipy.sparse.csr_matrix(F.Loff[hrows][:, loc[order]]) if hrows.size else scipy.sparse.csr_matrix((0, n))
    ident = np.arange(n + hrows.size, dtype=np.int64)
    gstart = np.searchsorted(gid_p, np.arange(n_groups + 1))
    fwd, bwd = [], []
    for g in range(n_groups):
        a, b = int(gstart[g]), int(gstart[g + 1])
        if a == b:
            continue
        Linv = scipy.linalg.solve_triangular(Lp[a:b, a:b].toarray(), np.eye(b - a), lower=True)
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
        fwd.append((_tiles_sparse(Lp[a:b, :a], a, ident) if a > 0 else [], inv_f))
        above = scipy.sparse.hstack([Lp[b:, a:b].T, Lh[:, a:b].T]).tocsr() if (b < n or hrows.size) else None
        bwd.append((_tiles_sparse(above, a, ident[b:]) if above is not None else [], inv_b))
    levels = []
    for (ta, inv) in reversed(bwd):
        levels.append(ta)
        levels.extend(inv)
    levels = [u for u in levels if u]
    n_bwd = len(levels)
    for (ta, inv) in fwd:
        levels.append(ta)
        levels.extend(inv)
    if hrows.size:
        levels.append(_tiles_sparse(Lh, n, ident))
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


def subdomain_plan(F, junc_face, d, NG, n_warps=RES_WARPS, groups=None):
    """
    F         : factor.Factor of the permuted cycle-space system
    junc_face : (Nj, 2) permuted faces of every junction, -1 none (CircuitTables.junc_face)
    d         : cut depth, P = 2^d subdomains; None: the subtrees of an ordering made with n_parts (F.blk_part)
    NG        : problem groups of 8 per chunk (PC = 8 * NG problems share one pass over the factor)
    groups    : tree heights after which a new level group starts (see _subdomain_levels); None = automatic
    """
    assert NG in (1, 2, 4, 8)
    n, nb = F.n, F.nb
    PC = 8 * NG
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
    top_rows = np.flatnonzero(row_sub < 0)
    n_top = int(top_rows.size)
    tix = np.full(n, -1, dtype=np.int64)
    tix[top_rows] = np.arange(n_top)

    # (subdomain, top row) couplings through the factor
    coo = F.Loff.tocoo()
    m = (row_sub[coo.row] < 0) & (row_sub[coo.col] >= 0)
    assert not np.any((row_sub[coo.row] >= 0) & (row_sub[coo.col] < 0)), "a local row precedes one of its separators"
    assert not np.any((row_sub[coo.row] >= 0) & (row_sub[coo.col] >= 0) & (row_sub[coo.row] != row_sub[coo.col])), \
        "two subdomains are coupled"
    coupled = [set() for _ in range(P)]
    subs_of_top = [[] for _ in range(n_top)]
    if m.any():
        pairs = np.unique(np.stack((row_sub[coo.col[m]], tix[coo.row[m]]), axis=1), axis=0)
        for s, k in pairs:
            coupled[int(s)].add(int(k))
            subs_of_top[int(k)].append(int(s))

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
                cand.update(subs_of_top[tix[jf[j, k]]])
        cand = sorted(cand) if cand else list(range(P))
        s = min(cand, key=lambda q: (load[q], q))
        owner[j] = s
        load[s] += 1
    for j in np.flatnonzero(has.any(axis=1)):
        for k in range(2):
            if has[j, k] and row_sub[jf[j, k]] < 0:
                coupled[int(owner[j])].add(int(tix[jf[j, k]]))

    plan = SubdomainPlan()
    plan.P, plan.NG, plan.PC, plan.d = P, NG, PC, d
    plan.n_top, plan.top_rows = n_top, top_rows.astype(np.int32)
    plan.n_top_pad = (n_top + 31) // 32 * 32
    if plan.n_top_pad > 8192:
        # the separators above the cut are solved by a DENSE inverse: 8192^2 float64 is 0.5 GB and 67 M multiply-adds
        # per problem and time step; beyond that the circuit needs a multi-level top (not built) and runs on the
        # streaming engine
        raise ValueError("subdomain plan: %d top rows are too many for the dense top product" % n_top)
    loc = [np.flatnonzero(row_sub == s) for s in range(P)]
    halo = [np.array(sorted(coupled[s]), dtype=np.int64) for s in range(P)]
    plan.n_loc = np.array([l.size for l in loc], dtype=np.int32)
    plan.n_halo = np.array([h.size for h in halo], dtype=np.int32)
    plan.n_loc_max = int(plan.n_loc.max())
    plan.n_rows = (int((plan.n_loc + plan.n_halo).max()) + 7) // 8 * 8 + 8
    if plan.n_rows * PC > 65536:
        raise ValueError("subdomain plan: %d rows x %d problems exceed the 16-bit element codes" % (plan.n_rows, PC))
    plan.hptr = np.concatenate(([0], np.cumsum(plan.n_halo))).astype(np.int32)
    plan.n_slots = int(plan.hptr[-1])
    plan.halo_top = (np.concatenate(halo) if plan.n_slots else np.zeros(0)).astype(np.int32)
    vrow = np.full((P, n), -1, dtype=np.int64)
    for s in range(P):
        vrow[s, loc[s]] = np.arange(loc[s].size)
        vrow[s, top_rows[halo[s]]] = loc[s].size + np.arange(halo[s].size)
    plan.vrow = vrow
    plan.row_sub = row_sub

    # ---- per-subdomain sweep programs: backward levels first (they open a time step), then forward
    plan.prog, plan.n_bwd = [], []
    plan.group_bounds = []
    smem_limit = 227 * 1024 - 20 * 1024                 # room for headers, cursors and the amplitude cache
    stage_cap = (smem_limit // 8 - plan.n_rows * PC) // (PC + 2)
    if stage_cap < 8:
        raise ValueError("subdomain plan: the right-hand sides leave no room for the staging rows")
    stage_cap = min(stage_cap, 4096)
    for s in range(P):
        levels, n_bwd, order, gb = _subdomain_levels(F, loc[s], top_rows[halo[s]], blk_of, stage_cap, groups)
        vrow[s, loc[s][order]] = np.arange(loc[s].size)
        plan.prog.append(_pack_levels(levels, NG, n_warps, n_bwd))
        plan.n_bwd.append(n_bwd)
        plan.group_bounds.append(gb)
    plan.n_bwd = np.asarray(plan.n_bwd, dtype=np.int32)
    plan.stage_rows = max(p["stage_rows"] for p in plan.prog)
    if n_top:
        # the staged top product keeps a ring of 4 stages x 8 KB (or 4 KB, 2 KB) in the staging rows
        for ring in (32768, 16384, 8192):
            need = -(-ring // ((PC + 2) * 8))
            if need <= stage_cap:
                plan.stage_rows = max(plan.stage_rows, need)
                break

    # ---- top: explicit inverse of the Schur complement, packed as FP64 MMA A fragments
    nTp = plan.n_top_pad
    if n_top:
        LTT = F.Lc[top_rows][:, top_rows].toarray()
        Linv = scipy.linalg.solve_triangular(LTT, np.eye(n_top), lower=True)
        Sinv = Linv.T @ Linv
        SP = np.zeros((nTp, nTp))
        SP[:n_top, :n_top] = Sinv
        plan.Sinv = Sinv
        plan.Sinv_packed = np.ascontiguousarray(SP.reshape(nTp // 8, 8, nTp // 4, 4).transpose(0, 2, 1, 3)).ravel()
    else:
        plan.Sinv = np.zeros((0, 0))
        plan.Sinv_packed = np.zeros(0)
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
    rows = np.where(has, vrow[owner[:, None], np.maximum(jf, 0)], -1)
    assert np.all(rows[has] >= 0)
    plan.junc_row = np.ascontiguousarray(rows[jorder].astype(np.int32))
    return plan


def face_tables(plan, face_ptr, face_junc, face_sign, junc_sign, c0):
    """Fixed-width per-(subdomain, row) lists of (device junction, sign / c0) for the face projection
    b = A (x / c0 - theta_s) (reference: time_evolution.py:560-569); a halo row lists only the junctions its
    subdomain owns, the partial sums meet in the top assembly."""
    P, n_rows = plan.P, plan.n_rows
    Nf = len(face_ptr) - 1
    g_of = np.repeat(np.arange(Nf), np.diff(face_ptr))
    j_of = np.asarray(face_junc, dtype=np.int64)
    r_of = plan.owner[j_of]
    row_of = plan.vrow[r_of, g_of]
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
    for s in range(P):
        g = np.flatnonzero(plan.row_sub == s)
        fidx[s, plan.vrow[s, g]] = g
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


def apply_subdomain_plan_host(plan, b_perm):
    """Solve through the plan on the host. b_perm: (Nf, PC) right-hand side in PERMUTED numbering."""
    P, NG = plan.P, plan.NG
    assert b_perm.shape[1] == plan.PC
    vec, z = [], []
    ctop = np.zeros((plan.n_slots, plan.PC))
    for s in range(P):
        v = np.zeros((plan.n_rows, plan.PC))
        loc = np.flatnonzero(plan.row_sub == s)
        v[plan.vrow[s, loc]] = b_perm[loc]
        ps = plan.prog[s]
        for l in range(plan.n_bwd[s], ps["n_levels"]):
            _run_level_host(ps, v, l, NG)
        ctop[plan.hptr[s]:plan.hptr[s + 1]] = v[plan.n_loc[s]: plan.n_loc[s] + plan.n_halo[s]]
        vec.append(v)
    rtop = b_perm[plan.top_rows].copy()
    for k in range(plan.n_top):
        for sl in plan.tslot[plan.tptr[k]:plan.tptr[k + 1]]:
            rtop[k] += ctop[sl]
    jtop = plan.Sinv @ rtop
    out = np.zeros_like(b_perm)
    out[plan.top_rows] = jtop
    for s in range(P):
        v = vec[s]
        v[plan.n_loc[s]: plan.n_loc[s] + plan.n_halo[s]] = jtop[plan.halo_top[plan.hptr[s]:plan.hptr[s + 1]]]
        ps = plan.prog[s]
        for l in range(plan.n_bwd[s]):
            _run_level_host(ps, v, l, NG)
        loc = np.flatnonzero(plan.row_sub == s)
        out[loc] = v[plan.vrow[s, loc]]
    return out
