"""
Geometric nested-dissection ordering of the cycle-space system matrix.

The per-step linear system  A (L + diag(1/(Cv+Rv))) A^T J = b  (reference: time_evolution.py:504-506)
has one unknown per face of a planar circuit, so the faces' centroids give a natural geometric
nested dissection: recursively cut the face set at the median of its longer extent, make the faces
on one side that touch the other side the separator, and recurse into the two halves.

The result is a block elimination tree: leaf blocks (small subdomains) and separator blocks. Blocks
at the same tree height are mutually independent, which is what the device triangular solves
exploit: the number of dependent phases is the tree height (~2 log2(sqrt(Nf)/leaf)), not the number
of rows.  (The reference leaves ordering to SuperLU's COLAMD: 386-14186 dependent row levels.)

Host setup code, runs once per circuit; vectorised over all subdomains of a tree depth.
"""
import numpy as np
import scipy.sparse

__all__ = ["nested_dissection"]


def nested_dissection(S, cx, cy, leaf_size=24, return_tree=False, n_parts=None, part_weights=None):
    """
    Parameters
    ----------
    S : (n, n) scipy sparse symmetric matrix (only its pattern is used)
    cx, cy : (n,) coordinates of the unknowns
    leaf_size : stop cutting when a subdomain has at most this many unknowns

    Returns
    -------
    perm : (n,) new-to-old permutation (row i of the permuted system is unknown perm[i])
    block_ptr : (nb + 1,) block b owns permuted rows block_ptr[b]:block_ptr[b+1]
    block_height : (nb,) 0 for leaves; a separator is one higher than the tallest block below it
    Blocks are numbered in post-order: every block comes after all blocks of its subtree.
    With return_tree=True also returns block_depth and block_dom: block b is the separator (or leaf) of
    the subdomain reached from the root by the cuts encoded in the binary digits of block_dom[b]
    (most significant digit = first cut); its subtree holds exactly the blocks (depth + t, dom * 2^t + r).

    n_parts : if given, the first cuts are made UNEVEN so that the tree has n_parts subtrees of about equal size
    (a domain that still has to yield p parts is cut p//2 : p - p//2); with return_tree=True a sixth array
    block_part is returned: the index of the part a block belongs to, or -1 for the separators above the parts.
    (The subdomain engine wants one part per (SM, problem chunk) pair, which is rarely a power of two.)
    part_weights : optional (n_parts,) relative sizes of the parts, in tree order (part k is the k-th leaf of the
    cut tree from left to right); used to balance the WORK of the parts rather than their row counts.
    """
    n = S.shape[0]
    cx = np.asarray(cx, dtype=np.double)
    cy = np.asarray(cy, dtype=np.double)
    Sc = scipy.sparse.coo_matrix(S)
    off = Sc.row != Sc.col
    ei, ej = Sc.row[off].astype(np.int64), Sc.col[off].astype(np.int64)
    xrank = np.unique(cx, return_inverse=True)[1].astype(np.int64)      # dense ranks of the coordinates
    yrank = np.unique(cy, return_inverse=True)[1].astype(np.int64)
    n_rank = int(max(xrank.max(initial=0), yrank.max(initial=0))) + 1
    dom = np.zeros(n, dtype=np.int64)          # subdomain id (path in the cut tree) of each unknown
    active = np.ones(n, dtype=bool)            # not yet placed in a block
    blk_depth = np.zeros(n, dtype=np.int64)    # (depth, dom) of the block each unknown ends up in
    depth, n_dom = 0, 1
    parts = np.array([max(1, int(n_parts or 1))], dtype=np.int64)      # parts still to be made out of each domain
    pos0 = np.zeros(1, dtype=np.int64)                                # tree-order position of a domain's first part
    wcum = np.concatenate(([0.0], np.cumsum(np.ones(int(parts[0])) if part_weights is None
                                              else np.asarray(part_weights, dtype=np.double))))
    assert wcum.size == parts[0] + 1
    part_of = np.full(n, -1, dtype=np.int64)
    if parts[0] == 1:
        part_of[:] = 0
    while True:
        idx = np.flatnonzero(active)
        if idx.size == 0:
            break
        size = np.bincount(dom[idx], minlength=n_dom)
        # a domain too small to be cut further becomes a single part
        small = (size <= leaf_size) & (parts > 1)
        for dm in np.flatnonzero(small):
            part_of[idx[dom[idx] == dm]] = pos0[dm]
            parts[dm] = 1
        leaf_nodes = idx[size[dom[idx]] <= leaf_size]
        blk_depth[leaf_nodes] = depth
        active[leaf_nodes] = False
        idx = np.flatnonzero(active)
        if idx.size == 0:
            break
        if depth >= 60:
            raise RuntimeError("nested dissection did not terminate")
        d = dom[idx]
        x, y = cx[idx], cy[idx]
        lo_x = np.full(n_dom, np.inf); hi_x = np.full(n_dom, -np.inf)
        lo_y = np.full(n_dom, np.inf); hi_y = np.full(n_dom, -np.inf)
        np.minimum.at(lo_x, d, x); np.maximum.at(hi_x, d, x)
        np.minimum.at(lo_y, d, y); np.maximum.at(hi_y, d, y)
        cut_x = (hi_x - lo_x) >= (hi_y - lo_y)          # cut across the longer extent
        coord = np.where(cut_x[d], x, y)
        other = np.where(cut_x[d], y, x)
        # (d, coord, other) lexicographically, as ONE integer sort: the coordinates enter through their dense ranks
        # (ties keep equal ranks, the sort is stable), which is several times faster than a three-key lexsort
        rc_ = np.where(cut_x[d], xrank[idx], yrank[idx])
        ro_ = np.where(cut_x[d], yrank[idx], xrank[idx])
        order = np.argsort((d * n_rank + rc_) * n_rank + ro_, kind="stable") if n_dom * n_rank * n_rank < 2 ** 62 \
            else np.lexsort((other, coord, d))
        sd = d[order]
        start = np.searchsorted(sd, np.arange(n_dom))
        rank = np.empty(idx.size, dtype=np.int64)
        rank[order] = np.arange(idx.size) - start[sd]
        half = np.zeros(n, dtype=np.int8)
        p_lo = parts // 2
        frac = (wcum[pos0 + p_lo] - wcum[pos0]) / np.maximum(wcum[pos0 + parts] - wcum[pos0], 1e-300)
        lo_count = np.where(parts > 1, np.rint(size * frac).astype(np.int64), (size + 1) // 2)
        # cut at a coordinate VALUE (that of the first unknown of the upper half): unknowns that share it stay on one
        # side, so on a lattice the cut runs along a lattice line, separators are straight and congruent regions get
        # congruent orderings; a domain whose lower half would end up empty is cut by rank instead
        first_hi = order[np.minimum(start + np.minimum(lo_count, np.maximum(size - 1, 0)), idx.size - 1)]
        cut_val = coord[first_hi]
        by_val = coord >= cut_val[d]
        n_lo = np.bincount(d, weights=~by_val, minlength=n_dom)
        use_val = (n_lo > 0)[d]
        half[idx] = np.where(use_val, by_val, rank >= lo_count[d])
        # separator: unknowns of the lower half coupled to the upper half of the same subdomain
        both = active[ei] & active[ej]
        e1, e2 = ei[both], ej[both]
        cross = (dom[e1] == dom[e2]) & (half[e1] == 0) & (half[e2] == 1)
        sep = np.unique(e1[cross])
        blk_depth[sep] = depth
        active[sep] = False
        idx = np.flatnonzero(active)
        dom[idx] = 2 * dom[idx] + half[idx]
        new_parts = np.ones(2 * n_dom, dtype=np.int64)
        new_parts[0::2] = np.where(parts > 1, p_lo, 1)
        new_parts[1::2] = np.where(parts > 1, parts - p_lo, 1)
        new_pos0 = np.zeros(2 * n_dom, dtype=np.int64)
        new_pos0[0::2] = pos0
        new_pos0[1::2] = np.where(parts > 1, pos0 + p_lo, pos0)
        fresh = np.zeros(2 * n_dom, dtype=bool)          # domains that just became a part of their own
        fresh[0::2] = (parts > 1) & (new_parts[0::2] == 1)
        fresh[1::2] = (parts > 1) & (new_parts[1::2] == 1)
        if fresh.any() and idx.size:
            sel = fresh[dom[idx]]
            part_of[idx[sel]] = new_pos0[dom[idx[sel]]]
        parts = new_parts
        pos0 = new_pos0
        n_dom *= 2
        depth += 1

    # A block is (depth, dom); its subtree covers cut-tree paths [dom << s, (dom << s) + 2^s - 1] with
    # s = maxd - depth. Post-order = by subtree end ascending, deeper first on ties.
    maxd = int(blk_depth.max()) if n else 0
    shift = maxd - blk_depth
    sub_hi = (dom << shift) + (np.int64(1) << shift) - 1
    keys, inv = np.unique(np.stack((sub_hi, -blk_depth), axis=1), axis=0, return_inverse=True)
    inv = inv.ravel()
    nb = keys.shape[0]
    # inside a block order geometrically along its longer extent (locality for row tiles)
    lo_x = np.full(nb, np.inf); hi_x = np.full(nb, -np.inf)
    lo_y = np.full(nb, np.inf); hi_y = np.full(nb, -np.inf)
    np.minimum.at(lo_x, inv, cx); np.maximum.at(hi_x, inv, cx)
    np.minimum.at(lo_y, inv, cy); np.maximum.at(hi_y, inv, cy)
    along_x = (hi_x - lo_x) >= (hi_y - lo_y)
    c1 = np.where(along_x[inv], cx, cy)
    c2 = np.where(along_x[inv], cy, cx)
    perm = np.lexsort((c2, c1, inv))
    block_ptr = np.concatenate(([0], np.cumsum(np.bincount(inv, minlength=nb))))
    b_depth = -keys[:, 1]
    b_hi = keys[:, 0]
    b_lo = b_hi - ((np.int64(1) << (maxd - b_depth)) - 1)
    height = np.zeros(nb, dtype=np.int64)
    stack = []
    for b in range(nb):                         # post-order: the stack top holds finished subtrees
        h = 0
        while stack and b_lo[stack[-1]] >= b_lo[b] and b_hi[stack[-1]] <= b_hi[b]:
            h = max(h, height[stack.pop()] + 1)
        height[b] = h
        stack.append(b)
    if return_tree:
        first = perm[block_ptr[:-1]] if nb else np.zeros(0, dtype=np.int64)
        if n_parts is not None:
            # parts are numbered by tree position; positions that stayed empty (tiny domains) are squeezed out
            bp = part_of[first]
            used = np.unique(bp[bp >= 0])
            if used.size and used.size != used[-1] + 1:
                remap = np.full(int(used[-1]) + 1, -1, dtype=np.int64)
                remap[used] = np.arange(used.size)
                bp = np.where(bp >= 0, remap[np.maximum(bp, 0)], -1)
            return perm, block_ptr, height, b_depth.astype(np.int64), dom[first], bp
        return perm, block_ptr, height, b_depth.astype(np.int64), dom[first]
    return perm, block_ptr, height
