"""
Host driver of the device time-evolution core.

``device_time_evolution_core(problem, th_store_mask, I_store_mask)`` has the contract of the
reference's ``time_evolution_core`` (reference: time_evolution.py:461-582): it returns ``th_out`` and
``I_out`` of shape (Nj, W, n_stored + 2) whose first two planes are the initial conditions
theta(-2), theta(-1) (and their supercurrents). Everything inside the reference's ``for`` loop runs
on the GPU through the C ABI of include/jjstep.h; the host only

  * builds the circuit tables and the compiled solve program once (``CircuitTables``),
  * classifies the four per-step inputs (sources.py) and uploads their tables chunk by chunk,
  * shards the problem axis over the requested devices (one host thread per GPU, no collective on
    the step path) and gathers the stored planes at the end.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import weakref

import numpy as np
import scipy.sparse

from . import _lib
from .current_phase_relation import harmonics
from .factor import factorize, streaming_program, system_matrix
from .subdomain import subdomain_plan, face_tables
from .sources import classify_source, nonnegative_factors, NotRankOne, ZERO, RANK1, DENSE

__all__ = ["device_time_evolution_core", "CircuitTables", "DeviceEngine", "last_run_stats"]

_TABLE_BYTES = 64 << 20      # budget for one chunk of rank-one amplitude tables
_DENSE_BYTES = 256 << 20     # budget for one chunk of dense tables / injected noise
_PLANE_BYTES = 48 << 30      # budget for the theta / current planes one jj_run call keeps on the device (180 GB HBM3e)
last_run_stats = {}          # filled by device_time_evolution_core (per device), for benchmarks and tests


class _PinnedBlock:
    """A page-locked host block exposed through the array interface; returns to the pool when the last array that
    views it is gone."""

    def __init__(self, pool, ptr, nbytes):
        self._pool, self._ptr, self._nbytes = pool, ptr, nbytes
        self.__array_interface__ = dict(data=(ptr, False), shape=(nbytes // 8,), typestr="<f8", version=3)

    def __del__(self):
        try:
            self._pool._release(self._ptr, self._nbytes)
        except Exception:
            pass


class _PinnedPool:
    """Result planes in page-locked memory (device -> host at PCIe speed, no page faults on fresh numpy pages). Blocks
    are reused across compute() calls of the same result size; pinning is slow, so only a few sizes are kept."""

    # JJ_PINNED_MAX_GB (default 8): page-locked memory in use + cached; results beyond that are ordinary numpy arrays
    MAX_PINNED = int(float(os.environ.get("JJ_PINNED_MAX_GB", "8")) * (1 << 30))
    MAX_BLOCK = MAX_PINNED // 2
    MAX_CACHED = MAX_PINNED // 2

    def __init__(self):
        self._free, self._cached, self._in_use, self._lock = {}, 0, 0, threading.Lock()

    def empty(self, shape, device=0):
        nbytes = int(np.prod(shape)) * 8
        if os.environ.get("JJ_PINNED_RESULTS", "1") == "0" or nbytes == 0 or nbytes > self.MAX_BLOCK:
            return np.empty(shape)
        evict = []
        with self._lock:
            lst = self._free.get(nbytes)
            ptr = lst.pop() if lst else None
            if ptr is not None:
                self._cached -= nbytes
            else:
                # no cached block of this size: cached blocks of other sizes give way before the result falls back to
                # pageable memory (a different problem size after a few others used to do that, at a fifth of the speed)
                for size in sorted(self._free, reverse=True):
                    while self._free[size] and self._in_use + self._cached + nbytes > self.MAX_PINNED:
                        evict.append(self._free[size].pop())
                        self._cached -= size
                if self._in_use + self._cached + nbytes > self.MAX_PINNED:
                    ptr = False
            if ptr is not False:
                self._in_use += nbytes
        for old in evict:
            _lib.load().jj_host_free(C.c_void_p(old))
        if ptr is False:
            return np.empty(shape)
        if ptr is None:
            out = C.c_void_p()
            if _lib.load().jj_host_alloc(int(device), nbytes, C.byref(out)) != 0 or not out.value:
                with self._lock:
                    self._in_use -= nbytes
                return np.empty(shape)
            ptr = out.value
        return np.asarray(_PinnedBlock(self, ptr, nbytes)).reshape(shape)

    def _release(self, ptr, nbytes):
        with self._lock:
            self._in_use -= nbytes
            if self._cached + nbytes <= self.MAX_CACHED and len(self._free.get(nbytes, ())) < 2:
                self._free.setdefault(nbytes, []).append(ptr)
                self._cached += nbytes
                return
        _lib.load().jj_host_free(C.c_void_p(ptr))


_pinned = _PinnedPool()


def _pinned_i32(shape, device=0):
    """int32 array of this shape in a pooled page-locked block (the pool hands out float64 blocks)"""
    n = int(np.prod(shape))
    return _pinned.empty(((n + 1) // 2,), device).view(np.int32)[:n].reshape(shape)


class CircuitTables:
    """
    Everything the device needs that depends only on (circuit, dt): coefficient vectors
    (reference: time_evolution.py:470-478), CSR form of A in the permuted face order, the two faces of
    every junction, and the compiled solve program of A (L + 1/(Cv+Rv)) A^T (reference: :504-506).
    """

    def __init__(self, circuit, dt, leaf_size=None, n_parts=None):
        A = scipy.sparse.csr_matrix(circuit.get_cycle_matrix())
        A.sum_duplicates()
        A.eliminate_zeros()
        A.sort_indices()
        Nf, Nj = A.shape
        self.Nj, self.Nf, self.dt = Nj, Nf, dt
        R = np.asarray(circuit._R(), dtype=np.double)
        Cc = np.asarray(circuit._C(), dtype=np.double)
        self.Rv = 1 / (dt * R)
        self.Cv = Cc / (dt ** 2)
        self.c0 = 1.0 * self.Rv + 1.0 * self.Cv
        self.c1 = -1.0 * self.Rv + -2.0 * self.Cv
        self.c2 = 0.0 * self.Rv + 1.0 * self.Cv
        self.Ic = np.ascontiguousarray(circuit._Ic(), dtype=np.double)
        self.n_parts = n_parts
        if leaf_size is None:
            # the subdomain engine applies explicit inverses of level groups: larger leaves give it fewer, fatter
            # tiles (8 is 4 % slower on cfg2, 36 and more 25 % slower: blocks beyond the 32 rows a single warp applies in
            # place). Re-measured on the round-2 kernel (profiles/r02_experiments.md): 16 beats the 32 of round 1 on
            # every named configuration (cfg2 88.8 vs 90.7 us per time step, cfg3 2300 vs 2326, cfg4 1490 vs 1526)
            leaf_size = int(os.environ.get("JJ_LEAF_SIZE", "16" if n_parts is not None else "8"))
        if Nf > 0:
            S = system_matrix(A, circuit._L(), self.Rv, self.Cv)
            if hasattr(circuit, "get_face_centroids"):
                cx, cy = circuit.get_face_centroids()
            else:
                cx, cy = _centroids_from_matrix(circuit, A)
            self.factor = factorize(S, cx, cy, leaf_size=leaf_size, n_parts=n_parts)
            # work balancing pays when a block owns ONE subdomain for the whole run and waits for the slowest block at
            # every step (cfg2: 18 subdomains); with many subdomains the blocks' items average out and every
            # balancing round would cost a full plan build
            if n_parts is not None and 1 < n_parts <= _BALANCE_MAX_PARTS and os.environ.get("JJ_SUB_BALANCE", "1") != "0":
                self.factor = self._balance_parts(A, S, cx, cy, leaf_size, n_parts, self.factor)
            self._program = None
            perm = self.factor.perm.astype(np.int64)
        else:
            self._program = None
            self.factor = None
            perm = np.zeros(0, dtype=np.int64)
        self._finish_tables(A, perm)

    @property
    def program(self):
        """Solve program of the streaming engine (built on first use: only dense per-step inputs need it)."""
        if self._program is None and self.factor is not None:
            self._program = streaming_program(self.factor)
        return self._program

    def _balance_parts(self, A, S, cx, cy, leaf_size, n_parts, F, rounds=None):
        """Re-run the dissection with part weights so that the WORK of the subdomains (sweep stream steps, junctions,
        rows; per-unit costs measured on B200) is even: every thread block waits for the slowest one at the grid
        barrier of each time step."""
        if rounds is None:
            rounds = int(os.environ.get("JJ_BAL_ROUNDS", "2"))
        weights = np.ones(n_parts)
        best, best_spread = F, None
        if A.shape[0] > 30000:
            rounds = 1              # every round is a full ordering + factorisation: seconds to minutes at this size
        for _ in range(rounds + 1):
            if F.blk_part is None or int(np.max(F.blk_part)) + 1 != n_parts:
                return best
            self._finish_tables(A, F.perm.astype(np.int64))
            try:
                plan = subdomain_plan(F, self.junc_face, None, 4)
            except ValueError:
                # subdomains too large for the shared-memory engine (big circuits): nothing to balance, the
                # streaming engine will run this circuit
                return best
            steps = np.array([p["n_steps"] for p in plan.prog], dtype=np.double)
            w_steps, w_junc, w_rows = (float(x) for x in os.environ.get("JJ_BAL_W", "41,52,42").split(","))
            cost = w_steps * steps + w_junc * np.diff(plan.junc_ptr) + w_rows * (plan.n_loc + plan.n_halo)
            spread = float(cost.max() / cost.mean())
            if best_spread is None or spread < best_spread:
                best, best_spread = F, spread
            if spread < 1.02 or _ == rounds:
                break
            weights = weights * (cost.mean() / cost) ** 0.8
            F = factorize(S, cx, cy, leaf_size=leaf_size, n_parts=n_parts, part_weights=weights)
        self.part_cost_spread = best_spread
        return best

    def _finish_tables(self, A, perm):
        Nf, Nj = A.shape
        self.perm = perm
        inv = np.empty(Nf, dtype=np.int64)
        inv[perm] = np.arange(Nf)
        self.inv_perm = inv
        Ap = A[perm] if Nf else A
        Ap.sort_indices()                       # ascending junction index inside a face, like the reference's CSC product
        self.face_ptr = Ap.indptr.astype(np.int32)
        self.face_junc = Ap.indices.astype(np.int32)
        self.face_sign = Ap.data.astype(np.int8)
        # the (at most two) faces of each junction, in ascending ORIGINAL face order like the reference's A^T product
        At = scipy.sparse.csr_matrix(A.T)
        At.sort_indices()
        cnt = np.diff(At.indptr)
        if cnt.size and cnt.max() > 2:
            raise ValueError("a junction borders more than two faces; the circuit is not a planar embedding")
        jf = np.full((Nj, 2), -1, dtype=np.int32)
        js = np.zeros((Nj, 2), dtype=np.int8)
        rows = np.repeat(np.arange(Nj), cnt)
        slot = np.arange(At.indices.size) - At.indptr[rows]
        jf[rows, slot] = inv[At.indices]
        js[rows, slot] = At.data
        self.junc_face, self.junc_sign = jf, js
        self._subdomain = {}

    # ------------------------------------------------------------------ subdomain engine plan
    SMEM_LIMIT = 226 * 1024          # dynamic shared memory of the step kernel (it has 1 KB of static shared memory)
    SMEM_LIMIT_HALF = 112 * 1024     # half blocks (8 warps), two to an SM

    def subdomain_smem_bytes(self, plan):
        PC = plan.PC
        aux = 3 * max(ps["n_levels"] * ps["n_warps"] + 1 for ps in plan.prog) + max(ps["n_levels"] for ps in plan.prog) \
            + 2 * max(len(ps["thdr"]) for ps in plan.prog) + 10
        return (plan.n_rows * PC + plan.stage_rows * (PC + 2)) * 8 + 64 * PC + 4 * aux

    def subdomain_plan(self, d, NG, n_chunks=1):
        n_warps = 8 if half_blocks() else 16
        key = (d, NG, n_chunks, os.environ.get("JJ_TT_MAX", ""), n_warps)
        if key not in self._subdomain:
            plan = subdomain_plan(self.factor, self.junc_face, d, NG, n_warps=n_warps, n_chunks=n_chunks)
            face_tables(plan, self.face_ptr, self.face_junc, self.face_sign, self.junc_sign, self.c0)
            self._subdomain[key] = plan
        return self._subdomain[key]

    def choose_subdomain(self, W, n_sm=148):
        """Pick (cut, problem groups NG, problem chunks) for the subdomain engine, or None. cut = None uses the n_parts subtrees
        this ordering was made with (see subdomain_layout); otherwise cut is a tree depth (2^cut subdomains).
        JJ_SUBDOMAIN="cut,NG" overrides ("p,NG" selects the parts)."""
        if self.factor is None:
            return None
        env = os.environ.get("JJ_SUBDOMAIN")
        if env:
            d, NG = env.split(",")
            NG = int(NG)
            return (None if d.strip().startswith("p") else int(d)), NG, -(-((W + 3) // 4 * 4) // (8 * NG))
        NG, chunks, _ = subdomain_layout(self.Nf, W, n_sm)
        if self.n_parts is not None and self.factor.blk_part is not None:
            cands = [None]
        else:
            max_d = int(self.factor.depth.max()) if self.factor.nb else 0
            d_hi = 0
            while d_hi < min(max_d - 1, 6) and (2 << d_hi) * chunks <= n_sm and self.Nf >= 64 * (2 << d_hi):
                d_hi += 1
            cands = [d for d in range(d_hi, 7) if d <= max(max_d - 1, 0)]
        for d in cands:
            try:
                plan = self.subdomain_plan(d, NG, chunks)
            except ValueError:
                continue
            if plan.prog[0]["n_warps"] == 8 and (plan.upper["n_fwd"] + plan.upper["n_bwd"] > 0):
                continue                     # half blocks have no upper program
            if self.subdomain_smem_bytes(plan) <= (self.SMEM_LIMIT_HALF if plan.prog[0]["n_warps"] == 8 else self.SMEM_LIMIT):
                return d, NG, chunks
        return None


_BALANCE_MAX_PARTS = 64
_ROWS_FIT = 560               # local rows of a subdomain that still fit in 227 KB at 32 problems with a ~15 % halo
_ROWS_MULTI = 520             # most local rows of a subdomain when a block works through several items per time step


def half_blocks():
    """JJ_SUB_HALF=1: blocks of 8 warps, two to an SM, each on its own (subdomain, chunk of 16 problems) item"""
    return os.environ.get("JJ_SUB_HALF", "0") == "1"


def subdomain_layout(Nf, W, n_sm=148):
    """(NG, chunks, n_parts) of the subdomain engine for W problems on a circuit with Nf faces: chunks of 8*NG
    problems, and as many subdomains as fill the SMs with (subdomain, chunk) thread blocks - but none smaller
    than ~45 faces (below that the top product and the barriers cost more than the local sweeps save; cfg1: 8 x 45
    faces run 6 % faster than 5 x 72)."""
    Wp = (W + 3) // 4 * 4
    NG = 4 if Wp > 16 else 2 if Wp > 8 else 1
    if half_blocks():
        # two resident blocks per SM: half the problems per chunk, twice the blocks
        NG = min(NG, 2)
        chunks = (Wp + 8 * NG - 1) // (8 * NG)
        n_parts = max(1, min(2 * n_sm // chunks, Nf // 45))
        if Nf <= _ROWS_FIT * n_parts:
            return NG, chunks, n_parts
    chunks = (Wp + 8 * NG - 1) // (8 * NG)
    n_parts = max(1, min(n_sm // chunks, Nf // 45))
    # larger circuits: the rows of a subdomain (local + halo) must fit in a block's shared memory, so there are more
    # (subdomain, chunk) items than SMs and every block loops over several items per time step
    if Nf > _ROWS_FIT * n_parts:
        # a power of two: every cut of the dissection is then a median cut, so a regular lattice falls into a handful
        # of congruent subdomain shapes whose sweep programs are shared (subdomain._subdomain_inputs): less host
        # setup, and the shared factor streams stay in L2
        n_parts = 1
        while Nf > _ROWS_MULTI * n_parts:
            n_parts *= 2
    return NG, chunks, n_parts


def _centroids_from_matrix(circuit, A):
    x, y = circuit.get_node_coordinates()
    n1, n2 = circuit.get_junction_nodes()
    jx, jy = 0.5 * (x[n1] + x[n2]), 0.5 * (y[n1] + y[n2])
    B = abs(A)
    deg = np.asarray(B.sum(axis=1)).ravel()
    return (B @ jx) / deg, (B @ jy) / deg


_tables_cache = {}
_engine_lock = threading.Lock()


def _tables_for(circuit, dt, n_parts=None):
    """Cache per circuit object, invalidated when component values change."""
    L = circuit._L()
    key = (id(circuit), float(dt), hash(np.asarray(circuit._R()).tobytes()), hash(np.asarray(circuit._C()).tobytes()),
           hash(np.asarray(circuit._Ic()).tobytes()), hash(L.data.tobytes()) ^ hash(L.indices.tobytes()),
           os.environ.get("JJ_LEAF_SIZE", ""), os.environ.get("JJ_TT_MAX", ""), n_parts)
    with _engine_lock:
        hit = _tables_cache.get(key)
        if hit is not None and hit[0]() is circuit:       # id() values are reused after garbage collection
            return hit[1]
    tab = CircuitTables(circuit, dt, n_parts=n_parts)
    try:
        ref = weakref.ref(circuit)
    except TypeError:
        ref = (lambda c=circuit: c)
    with _engine_lock:
        # a few entries: an annealing schedule alternates between dt and dt / 2 on the same circuit
        while len(_tables_cache) >= 4:
            _tables_cache.pop(next(iter(_tables_cache)))
        _tables_cache[key] = (ref, tab)
    return tab


class DeviceEngine:
    """One GPU, one shard of the problem axis. Thin object wrapper over the C ABI."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        rc = self.lib.jj_create(int(device), C.byref(self.h))
        if rc != 0:
            raise RuntimeError("jj_create failed: " + self.lib.jj_last_error(None).decode())
        self.device = device
        self._keep = []

    def close(self):
        if self.h:
            self.lib.jj_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.jj_last_error(self.h).decode()
            if rc == _lib.JJ_ENONFINITE:
                raise FloatingPointError(msg)
            raise (ValueError if rc == -2 else RuntimeError)(f"libjjstep error {rc}: {msg}")

    # --- setup -----------------------------------------------------------------------------
    def set_circuit(self, tab: CircuitTables, cpr, with_program=True):
        a, b = harmonics(cpr)
        a, b = _lib.c_f64(a), _lib.c_f64(b)
        c = _lib.JJCircuit()
        c.Nj, c.Nf = tab.Nj, tab.Nf
        keep = [tab.face_ptr, tab.face_junc, tab.face_sign, np.ascontiguousarray(tab.junc_face),
                np.ascontiguousarray(tab.junc_sign), tab.Ic, _lib.c_f64(tab.c0), _lib.c_f64(tab.c1),
                _lib.c_f64(tab.c2), a, b]
        c.face_ptr, c.face_junc, c.face_sign = _lib.i32(keep[0]), _lib.i32(keep[1]), _lib.i8(keep[2])
        c.junc_face, c.junc_sign = _lib.i32(keep[3]), _lib.i8(keep[4])
        c.Ic, c.c0, c.c1, c.c2 = _lib.f64(keep[5]), _lib.f64(keep[6]), _lib.f64(keep[7]), _lib.f64(keep[8])
        c.cpr_harmonics = len(a) - 1
        c.cpr_a, c.cpr_b = _lib.f64(a), _lib.f64(b)
        self._ck(self.lib.jj_set_circuit(self.h, C.byref(c)))
        self.tab = tab
        self.has_streaming_program = False
        self.set_streaming_program(with_program)

    def set_streaming_program(self, with_program=True):
        """Upload the solve program of the streaming engine (with_program=False: empty sweeps, for engines that only
        ever run the subdomain kernel; the program of a large circuit takes long to build and is rarely needed)."""
        tab = self.tab
        prog = tab.program if (with_program and tab.factor is not None) else None
        # (the structs hold raw pointers into these arrays: the dictionaries must outlive the call)
        held = [prog.sweeps[name] if prog is not None else _empty_sweep() for name in ("fwd", "bwd")]
        sweeps = [self._sweep_struct(sw) for sw in held]
        self._ck(self.lib.jj_set_solver(self.h, C.byref(sweeps[0]), C.byref(sweeps[1])))
        del held
        self.has_streaming_program = prog is not None or tab.factor is None

    @staticmethod
    def _sweep_struct(sw):
        s = _lib.JJSweep()
        s.n_levels = len(sw["level_ptr"]) - 1
        s.level_ptr, s.group_ptr = _lib.i32(sw["level_ptr"]), _lib.i32(sw["group_ptr"])
        s.n_tiles = len(sw["tile_row0"])
        s.tile_row0, s.tile_nrows = _lib.i32(sw["tile_row0"]), _lib.i32(sw["tile_nrows"])
        s.tile_lpr, s.tile_nsteps = _lib.i32(sw["tile_lpr"]), _lib.i32(sw["tile_nsteps"])
        s.tile_flags = _lib.i32(sw["tile_flags"])
        s.tile_col_off, s.tile_val_off = _lib.i64(sw["tile_col_off"]), _lib.i64(sw["tile_val_off"])
        s.n_cols, s.cols = sw["cols"].size, _lib.i32(sw["cols"])
        s.n_vals, s.vals = sw["vals"].size, _lib.f64(sw["vals"])
        s.stage_rows = sw["stage_rows"]
        s.tile_stage_off = _lib.i32(sw["tile_stage_off"])
        return s

    def set_subdomain(self, d, NG, n_chunks=1):
        """Upload the subdomain-engine plan for cut depth d (None: the n_parts subtrees of the ordering) and NG
        groups of 8 problems per chunk (n_chunks: the chunk count the upper program is sized for)."""
        plan = self.tab.subdomain_plan(d, NG, n_chunks)
        p = _lib.JJSubdomainPlan()
        p.P, p.NG, p.n_rows, p.n_loc_max, p.stage_rows = plan.P, plan.NG, plan.n_rows, plan.n_loc_max, plan.stage_rows
        p.n_top, p.n_up_pad, p.n_slots = plan.n_top, plan.n_up_pad, plan.n_slots
        p.tt0, p.n_tt, p.n_tt_pad = plan.tt0, plan.n_tt, plan.n_tt_pad
        keep = []

        def a32(x):
            x = np.ascontiguousarray(x, dtype=np.int32)
            if x.size == 0:
                x = np.zeros(1, dtype=np.int32)
            keep.append(x)
            return _lib.i32(x)

        def a64f(x):
            x = np.ascontiguousarray(x, dtype=np.double)
            if x.size == 0:
                x = np.zeros(1)
            keep.append(x)
            return _lib.f64(x)
        p.n_loc, p.n_halo, p.hptr, p.halo_top = a32(plan.n_loc), a32(plan.n_halo), a32(plan.hptr), a32(plan.halo_top)
        p.tptr, p.tslot, p.top_face = a32(plan.tptr), a32(plan.tslot), a32(plan.top_rows)
        p.Sinv_packed = a64f(plan.Sinv_packed)
        up = plan.upper
        p.up_RB, p.up_KB, p.n_up_fwd, p.n_up_bwd = up["RB"], up["KB"], up["n_fwd"], up["n_bwd"]
        p.n_up_tasks = len(up["task"])
        p.up_phase_ptr, p.up_phase_split, p.up_task = a32(up["phase_ptr"]), a32(up["phase_split"]), a32(up["task"])
        aoff = np.ascontiguousarray(up["task_aoff"] if len(up["task_aoff"]) else np.zeros(1), dtype=np.int64)
        keep.append(aoff)
        p.up_task_aoff = _lib.i64(aoff)
        p.n_up_cols, p.up_cols = up["cols"].size, a32(up["cols"])
        p.n_up_vals, p.up_A = up["A"].size, a64f(up["A"])

        memo = {}

        def sub_prog(ps, n_bwd):
            # congruent subdomains share one program object: hand the library the same host pointers, it uploads them once
            if id(ps) in memo:
                return memo[id(ps)]
            r = memo[id(ps)] = _lib.JJSubProgram()
            r.n_levels, r.n_bwd, r.n_warps, r.n_tiles = ps["n_levels"], int(n_bwd), ps["n_warps"], len(ps["thdr"])
            r.wt_ptr, r.ws_ptr, r.thdr, r.lstaged = a32(ps["wt_ptr"]), a32(ps["ws_ptr"]), a32(ps["thdr"]), a32(ps["lstaged"])
            r.n_steps = ps["n_steps"]
            st = ps["stream"] if ps["stream"].size else np.zeros(1, dtype=np.uint8)
            keep.append(st)
            r.stream = st.ctypes.data_as(C.POINTER(C.c_uint8))
            return r
        progs = (_lib.JJSubProgram * plan.P)(*[sub_prog(ps, nb) for ps, nb in zip(plan.prog, plan.n_bwd)])
        p.prog = progs
        p.junc_ptr, p.junc_orig, p.junc_row = a32(plan.junc_ptr), a32(plan.junc_orig), a32(plan.junc_row)
        js = np.ascontiguousarray(plan.junc_sign, dtype=np.int8)
        keep.append(js)
        p.junc_sign = _lib.i8(js)
        p.face_K = plan.face_K
        p.face_ell_j, p.face_ell_c, p.face_fidx = a32(plan.face_ell_j), a64f(plan.face_ell_c), a32(plan.face_fidx)
        self._ck(self.lib.jj_set_subdomain_plan(self.h, C.byref(p)))
        self.subdomain_config = (d, NG, n_chunks)

    def debug_subdomain_solve(self, b):
        bp = _lib.c_f64(np.asarray(b)[self.tab.perm])
        Jp = np.empty_like(bp)
        self._ck(self.lib.jj_debug_subdomain_solve(self.h, _lib.f64(bp), _lib.f64(Jp)))
        J = np.empty_like(Jp)
        J[self.tab.perm] = Jp
        return J

    def set_problem(self, W, dt, seed=0, problem_offset=0, engine=_lib.JJ_ENGINE_AUTO):
        self.W = W
        self._ck(self.lib.jj_set_problem(self.h, W, float(dt), int(seed) & (2 ** 64 - 1), int(problem_offset), engine))

    def set_engine(self, engine):
        """Switch the step engine of the current problem (the state stays on the device)."""
        self._ck(self.lib.jj_set_engine(self.h, int(engine)))

    def set_state(self, th_m1, th_m2):
        a, b = _lib.c_f64(th_m1), _lib.c_f64(th_m2)
        assert a.shape == (self.tab.Nj, self.W) and b.shape == (self.tab.Nj, self.W)
        self._ck(self.lib.jj_set_state(self.h, _lib.f64(a), _lib.f64(b)))

    def get_state(self, previous=True, out=None):
        """(theta(-1), theta(-2)) of the device state; previous=False fetches theta(-1) only (-> (theta(-1), None));
        out: a C-contiguous (Nj, W) float64 array to receive theta(-1) (e.g. page-locked)"""
        ok = out is not None and out.shape == (self.tab.Nj, self.W) and out.dtype == np.double and out.flags.c_contiguous
        a = out if ok else np.empty((self.tab.Nj, self.W))
        b = np.empty((self.tab.Nj, self.W)) if previous else None
        self._ck(self.lib.jj_get_state(self.h, _lib.f64(a), _lib.f64(b) if previous else None))
        return a, b

    def set_source(self, which, kind, is_static, base=None):
        basep = _lib.f64(_lib.c_f64(base)) if base is not None else None
        self._ck(self.lib.jj_set_source(self.h, which, kind, int(is_static), basep))

    def upload_source(self, which, i0, table):
        t = _lib.c_f64(table)
        self._ck(self.lib.jj_upload_source(self.h, which, int(i0), int(t.shape[0]), _lib.f64(t)))

    def upload_noise(self, i0, Z):
        if Z is None:
            self._ck(self.lib.jj_upload_noise(self.h, 0, 0, None))
            return
        z = _lib.c_f64(Z)
        self._ck(self.lib.jj_upload_noise(self.h, int(i0), int(z.shape[0]), _lib.f64(z)))

    def alloc_outputs(self, n_th, n_I):
        self._ck(self.lib.jj_alloc_outputs(self.h, int(n_th), int(n_I)))

    def run(self, i0, n, th_plane=None, I_plane=None):
        tp = np.ascontiguousarray(th_plane if th_plane is not None else -np.ones(n), dtype=np.int64)
        ip = np.ascontiguousarray(I_plane if I_plane is not None else -np.ones(n), dtype=np.int64)
        self._ck(self.lib.jj_run(self.h, int(i0), int(n), _lib.i64(tp), _lib.i64(ip)))

    def _fetch_target(self, n, out):
        if out is not None and out.shape == (n, self.tab.Nj, self.W) and out.flags.c_contiguous \
                and out.dtype == np.double and out.flags.writeable:
            return out, True
        return np.empty((n, self.tab.Nj, self.W)), False

    def fetch_theta(self, p0, n, out=None):
        """Stored theta planes [p0, p0 + n) as (n, Nj, W); written straight into ``out`` when that is a contiguous
        array of this shape (saves one pass over the result on the host)."""
        buf, direct = self._fetch_target(n, out)
        self._ck(self.lib.jj_fetch_theta(self.h, int(p0), int(n), _lib.f64(buf)))
        if out is not None and not direct:
            out[...] = buf
        return buf

    def fetch_current(self, p0, n, out=None):
        buf, direct = self._fetch_target(n, out)
        self._ck(self.lib.jj_fetch_current(self.h, int(p0), int(n), _lib.f64(buf)))
        if out is not None and not direct:
            out[...] = buf
        return buf

    def debug_noise(self, step):
        out = np.empty((self.tab.Nj, self.W))
        self._ck(self.lib.jj_debug_noise(self.h, int(step), _lib.f64(out)))
        return out

    def debug_solve(self, b):
        """Solve S J = b on the device; b, J are (Nf, W) in ORIGINAL face numbering."""
        bp = _lib.c_f64(np.asarray(b)[self.tab.perm])
        Jp = np.empty_like(bp)
        self._ck(self.lib.jj_debug_solve(self.h, _lib.f64(bp), _lib.f64(Jp)))
        J = np.empty_like(Jp)
        J[self.tab.perm] = Jp
        return J

    # --- annealing support (reference: time_evolution.py:1142-1191) ------------------------------
    def restart_at_rest(self):
        """theta(-2) := theta(-1) on the device (reference: time_evolution.py:1169-1171)."""
        self._ck(self.lib.jj_restart_at_rest(self.h))

    def adopt_state_at_rest(self, other):
        """theta(-1) = theta(-2) := theta(-1) of another engine on the same device, device to device."""
        self._ck(self.lib.jj_adopt_state_at_rest(self.h, other.h))

    def vortex_configuration(self, plane=-1):
        """n = -A round(theta / 2 pi) of a stored theta plane (-1: the current state), (Nf, W) int array in the
        ORIGINAL face numbering (reference: time_evolution.py:734-755)."""
        out = np.zeros((self.tab.Nf, self.W), dtype=np.int32)
        self._ck(self.lib.jj_vortex_configuration(self.h, int(plane), _lib.i32(out)))
        n = np.empty_like(out)
        n[self.tab.perm] = out
        return n.astype(int)

    def vortex_configurations(self, plane0, n_planes):
        """n of the stored theta planes [plane0, plane0 + n_planes) in one call: (n_planes, Nf, W) int32, ORIGINAL face
        numbering; the theta planes themselves stay on the device."""
        out = _pinned_i32((int(n_planes), self.tab.Nf, self.W), self.device)
        if n_planes and self.tab.Nf:
            order = np.ascontiguousarray(self.tab.perm, dtype=np.int32)     # permuted face p is face perm[p] of the circuit
            self._ck(self.lib.jj_vortex_configurations(self.h, int(plane0), int(n_planes), _lib.i32(out), _lib.i32(order)))
        else:
            out[...] = 0
        return out

    def observe_begin(self, first_step, interval):
        """Running observables: every ``interval``-th step from ``first_step`` on, the step kernel adds the vortex
        configuration into per-(face, problem) sums and keeps the phases of the first and the latest observation
        (0 switches it off)."""
        self._ck(self.lib.jj_observe_begin(self.h, int(first_step), int(interval)))

    def observe_fetch(self, marks=True):
        """-> (count, nsum (Nf, W) int32 in the ORIGINAL face numbering, theta_first, theta_latest ((Nj, W) or None))"""
        cnt = C.c_int64(0)
        ns = _pinned_i32((self.tab.Nf, self.W), self.device)
        if self.tab.Nf == 0:
            ns[...] = 0
        t0 = _pinned.empty((self.tab.Nj, self.W), self.device) if marks else None
        t1 = _pinned.empty((self.tab.Nj, self.W), self.device) if marks else None
        order = np.ascontiguousarray(self.tab.perm, dtype=np.int32)
        self._ck(self.lib.jj_observe_fetch(self.h, C.byref(cnt), _lib.i32(ns), _lib.f64(t0) if marks else None,
                                           _lib.f64(t1) if marks else None, _lib.i32(order)))
        return int(cnt.value), ns, t0, t1

    def anneal(self, first_interval, n_intervals, steps, upper, T_factor, norm, T):
        """The temperature schedule of n_intervals intervals on the device (jj_anneal): -> (T after the last interval,
        profiles (n_intervals, W), device milliseconds). upper: the mobility target of every interval."""
        T = np.ascontiguousarray(T, dtype=np.double).copy()
        up = np.ascontiguousarray(upper, dtype=np.double)
        assert up.shape == (n_intervals,) and T.shape == (self.W,)
        prof = np.zeros((n_intervals, self.W))
        ms = C.c_double(0.0)
        self._ck(self.lib.jj_anneal(self.h, int(first_interval), int(n_intervals), int(steps), _lib.f64(up), float(T_factor),
                                    1 / T_factor, float(norm), _lib.f64(T), _lib.f64(prof), C.byref(ms)))
        return T, prof, float(ms.value)

    def vortex_mobility_sums(self, plane0, n_planes):
        """(W,) integer sums over faces and consecutive stored planes of |n(t+1) - n(t)|
        (numerator of the reference's get_vortex_mobility, time_evolution.py:1128-1133)."""
        out = np.zeros(self.W, dtype=np.int64)
        self._ck(self.lib.jj_vortex_mobility(self.h, int(plane0), int(n_planes), _lib.i64(out)))
        return out

    def stats(self):
        s = _lib.JJStats()
        self._ck(self.lib.jj_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}


def _empty_sweep():
    z32, z64 = np.zeros(1, np.int32), np.zeros(1, np.int64)
    return dict(level_ptr=z32, group_ptr=z32, tile_row0=z32[:0], tile_nrows=z32[:0], tile_lpr=z32[:0],
                tile_nsteps=z32[:0], tile_flags=z32[:0], tile_stage_off=z32[:0], tile_col_off=z64[:0], tile_val_off=z64[:0],
                cols=z32[:0], vals=np.zeros(0), stage_rows=0)


# ----------------------------------------------------------------------------------------------
class _ShardInputs:
    """The four classified inputs restricted to problems [w0, w1)."""

    def __init__(self, specs, w0, w1):
        self.specs, self.w0, self.w1 = specs, w0, w1

    def amp(self, name, i0, i1):
        return self.specs[name].amp_chunk(i0, i1)[:, self.w0:self.w1]

    def dense(self, name, i0, i1):
        return self.specs[name].dense_chunk(i0, i1)[:, :, self.w0:self.w1]


def _classify_all(problem, tab):
    Nj, Nf, W, Nt = tab.Nj, tab.Nf, problem.get_problem_count(), problem._Nt()
    raw = getattr(problem, "_raw_sources", None)
    if raw is None:
        # a reference-style problem object (the reference's own TimeEvolutionProblem bound to this core):
        # its inputs are callables or (N, W, Nt) broadcast views, with the time-dependence flags frozen at construction;
        # an input that is not time dependent is read at step 0 only (reference: time_evolution.py:509-531)
        raw = {}
        for name, attr, flag in (("f", "external_flux", "_f_is_timedep"), ("Is", "current_sources", "_Is_is_timedep"),
                                 ("Vs", "voltage_sources", "_Vs_is_timedep"), ("T", "temperature", "_T_is_timedep")):
            cur = getattr(problem, attr)
            if not getattr(problem, flag, True):
                cur = np.asarray(cur(0))[..., None] if callable(cur) else np.asarray(cur)[:, :, 0:1]
            raw[name] = cur
    else:
        # an input attribute replaced after construction (the reference's annealing loop assigns prob.temperature
        # between compute() calls, time_evolution.py:1166): the reference reads the new attribute, with the
        # time-dependence flag frozen at construction - a "constant" input is sliced at step 0 once (:509-519)
        raw = dict(raw)
        kept = getattr(problem, "_kept_sources", {})
        for name, attr, flag in (("f", "external_flux", "_f_is_timedep"), ("Is", "current_sources", "_Is_is_timedep"),
                                 ("Vs", "voltage_sources", "_Vs_is_timedep"), ("T", "temperature", "_T_is_timedep")):
            cur = getattr(problem, attr)
            if name in kept and cur is not kept[name]:
                if not getattr(problem, flag):
                    cur = np.asarray(cur(0))[..., None] if callable(cur) else np.asarray(cur)[:, :, 0:1]
                raw[name] = cur
    specs = dict(Is=classify_source(raw["Is"], Nj, W, Nt), f=classify_source(raw["f"], Nf, W, Nt),
                 Vs=classify_source(raw["Vs"], Nj, W, Nt),
                 T=nonnegative_factors(classify_source(raw["T"], Nj, W, Nt)))
    return specs


def resolve_noise_seed(problem, noisy=True):
    """Philox seed of one compute() call. An explicit ``noise_seed`` is used as it is (reproducible runs); without
    one every call draws a fresh 64-bit seed from numpy's global generator - like the reference, whose draws advance
    ``np.random`` from call to call (time_evolution.py:533-538), so repeated runs, runs continued in segments and
    annealing intervals see independent noise, and ``np.random.seed`` still makes a script reproducible. Nothing is
    drawn when the temperature is zero (the reference draws nothing then either)."""
    seed = getattr(problem, "noise_seed", None)
    if seed is not None:
        return int(seed)
    if not noisy:
        return 0
    return int(np.random.randint(0, np.iinfo(np.int64).max, dtype=np.int64))


def _chunk_length(specs, tab, W, Nt, has_replay):
    K = Nt
    for name, s in specs.items():
        if s.kind == ZERO or (s.static and not (name == "Vs" and s.kind == RANK1)):
            continue
        N = tab.Nf if name == "f" else tab.Nj
        per = W * 8 if s.kind == RANK1 else N * W * 8
        K = min(K, max(1, (_TABLE_BYTES if s.kind == RANK1 else _DENSE_BYTES) // per))
    if has_replay:
        K = min(K, max(1, _DENSE_BYTES // (tab.Nj * W * 8)))
    return max(1, K)


_WHICH = dict(Is=_lib.JJ_SRC_IS, f=_lib.JJ_SRC_F, Vs=_lib.JJ_SRC_VS, T=_lib.JJ_SRC_T)


_engine_cache = {}           # device -> [(key, DeviceEngine)]: circuit, solver and plan stay uploaded between compute() calls
_ENGINES_PER_DEVICE = 2


def _engine_for(tab, cpr, dev, W, engine_kind, dense_inputs=False):
    """(key, DeviceEngine, engine kind to run) with the circuit tables and the plan of the engine that applies
    uploaded. Repeated compute() calls on the same circuit (annealing loops, parameter sweeps) reuse it: only the
    problem state is re-created. The subdomain engine runs whenever a plan fits and no input is a dense per-step
    table; the streaming engine (and its solve program, built on first use) covers the rest."""
    _lib.load()              # fails loudly when the CUDA library is missing, cached engine or not
    a, b = harmonics(cpr)
    sub_cfg = None
    if engine_kind in (_lib.JJ_ENGINE_AUTO, _lib.JJ_ENGINE_SUBDOMAIN) and not dense_inputs:
        sub_cfg = tab.choose_subdomain(W)
    if sub_cfg is None and engine_kind == _lib.JJ_ENGINE_SUBDOMAIN:
        raise ValueError("subdomain engine requested but " + ("an input is a dense per-step table" if dense_inputs
                                                              else "no plan fits in shared memory"))
    run_kind = _lib.JJ_ENGINE_SUBDOMAIN if sub_cfg is not None else _lib.JJ_ENGINE_STREAMING
    key = (id(tab), tuple(a), tuple(b), sub_cfg)
    eng = None
    with _engine_lock:
        entries = _engine_cache.setdefault(dev, [])
        for i, (k, e) in enumerate(entries):
            if k == key:
                del entries[i]
                eng = e
                break
    if eng is None:
        eng = DeviceEngine(dev)
        eng.set_circuit(tab, cpr, with_program=sub_cfg is None)
        if sub_cfg is not None:
            eng.set_subdomain(*sub_cfg)
    elif run_kind == _lib.JJ_ENGINE_STREAMING and not eng.has_streaming_program:
        eng.set_streaming_program(True)
    return key, eng, run_kind


def _release_engine(dev, key, eng, ok):
    if not ok:
        eng.close()
        return
    # two engines per device stay uploaded: an annealing schedule alternates between the tables of dt and dt / 2
    with _engine_lock:
        entries = _engine_cache.setdefault(dev, [])
        entries.append((key, eng))
        old = [entries.pop(0)] if len(entries) > _ENGINES_PER_DEVICE else []
    for _, e in old:
        e.close()


def _setup_sources(eng, specs, sh, tab):
    """Declare the device form of each input and upload the parts that do not change with time."""
    for name, s in specs.items():
        which = _WHICH[name]
        if s.kind == ZERO:
            eng.set_source(which, _lib.JJ_KIND_ZERO, True)
            continue
        if s.kind == RANK1:
            base = s.base
            if name == "T":
                base = np.sqrt(2.0 * s.base * tab.Rv)
            elif name == "f":
                base = s.base[tab.perm]
            static = s.static and name != "Vs"
            eng.set_source(which, _lib.JJ_KIND_RANK1, static, base)
            if static:
                amp = sh.amp(name, 0, 1)
                eng.upload_source(which, 0, np.sqrt(amp) if name == "T" else amp)
        else:
            eng.set_source(which, _lib.JJ_KIND_DENSE, s.static)
            if s.static:
                eng.upload_source(which, 0, _dense_for_device(name, sh.dense(name, 0, 1), tab))


def _run_shard(problem, tab, specs, dev, w0, w1, th_mask, I_mask, th_host, I_host, engine_kind, stats_out, seed=0,
               extras=None):
    """Integrate problems [w0, w1) on one device and write stored planes into th_host / I_host
    (plane-major (n_planes + 2, Nj, W) arrays, planes 0 and 1 are the initial conditions)."""
    Nj, Nf, Nt, dt = tab.Nj, tab.Nf, problem._Nt(), problem._dt()
    W = w1 - w0
    dense = any(s.kind == DENSE for s in specs.values())
    key, eng, engine_kind = _engine_for(tab, problem.current_phase_relation, dev, W, engine_kind, dense)
    ok = False
    try:
        eng.set_problem(W, dt, seed=seed, problem_offset=w0, engine=engine_kind)
        at_rest = getattr(problem, "starts_at_rest_with_zero_phases", None)
        if not (at_rest is not None and at_rest()):       # (jj_set_problem has cleared the state: nothing to upload)
            pair = getattr(problem, "initial_phases", None)
            m1, m2 = pair() if pair is not None else (problem.config_at_minus_1, problem.config_at_minus_2)
            eng.set_state(m1[:, w0:w1], m2[:, w0:w1])
        ex = extras or {}
        lead = ex.get("lead", 2)
        if ex.get("interval"):
            eng.observe_begin(ex.get("first", 0), ex["interval"])
        fetch_theta = ex.get("fetch_theta", True)
        sh = _ShardInputs(specs, w0, w1)
        replay = getattr(problem, "noise_replay", None)
        # static parts
        vs_cum = np.zeros(W)          # running sum of Vs amplitude * dt (rank-one voltage sources)
        _setup_sources(eng, specs, sh, tab)
        K = _chunk_length(specs, tab, W, Nt, replay is not None)
        th_idx = np.cumsum(th_mask) - 1      # plane index (without the +2 offset) of each stored step
        I_idx = np.cumsum(I_mask) - 1
        total_ms = 0.0
        i0 = 0
        # planes kept on the device per jj_run call: a chunk of steps ends before its stored planes outgrow the budget
        kept = np.concatenate(([0], np.cumsum(th_mask.astype(np.int64) + I_mask.astype(np.int64))))
        plane_budget = max(1, int(_PLANE_BYTES // (Nj * W * 8)))
        while i0 < Nt:
            i1 = min(Nt, i0 + K)
            if kept[i1] - kept[i0] > plane_budget:
                i1 = max(i0 + 1, int(np.searchsorted(kept, kept[i0] + plane_budget, side="right")) - 1)
            n = i1 - i0
            try:
                tables = []
                for name, s in specs.items():
                    if s.kind == ZERO:
                        continue
                    if s.kind == RANK1:
                        if name == "Vs":
                            amp = sh.amp(name, i0, i1)
                            if s.static:
                                amp = np.broadcast_to(amp, (n, W))
                            cum = vs_cum[None, :] + np.concatenate((np.zeros((1, W)), np.cumsum(amp * dt, axis=0)[:-1]), axis=0)
                            tables.append((name, cum, np.sum(amp * dt, axis=0)))
                        elif not s.static:
                            amp = sh.amp(name, i0, i1)
                            tables.append((name, np.sqrt(amp) if name == "T" else amp, None))
                    elif not s.static:
                        tables.append((name, _dense_for_device(name, sh.dense(name, i0, i1), tab), None))
            except NotRankOne:
                # a callable that was base[e] * amp(i)[w] at the probed steps is not any more: from this chunk of steps
                # on it is uploaded as dense per-step tables, which only the streaming engine reads (the state stays
                # where it is: both engines hand it over through the canonical arrays)
                name = _first_not_rank_one(specs, sh, i0, i1)
                if name == "Vs":
                    raise NotImplementedError("a callable voltage source stopped being base[e] * amp(i)[w] during the "
                                              "run; pass it as an array") from None
                specs = dict(specs)
                specs[name] = specs[name].as_dense()
                sh = _ShardInputs(specs, w0, w1)
                if engine_kind != _lib.JJ_ENGINE_STREAMING:
                    engine_kind = _lib.JJ_ENGINE_STREAMING
                    if not eng.has_streaming_program:
                        eng.set_streaming_program(True)
                    eng.set_engine(engine_kind)
                eng.set_source(_WHICH[name], _lib.JJ_KIND_DENSE, False)
                K = _chunk_length(specs, tab, W, Nt, replay is not None)
                continue
            for name, table, vs_add in tables:
                eng.upload_source(_WHICH[name], i0, table)
                if vs_add is not None:
                    vs_cum = vs_cum + vs_add
            if replay is not None and specs["T"].kind != ZERO:
                if callable(replay):
                    Z = np.stack([np.asarray(replay(i))[:, w0:w1] for i in range(i0, i1)])
                else:
                    Z = np.asarray(replay)[i0:i1, :, w0:w1]
                eng.upload_noise(i0, Z)
            tm, im = th_mask[i0:i1], I_mask[i0:i1]
            n_th, n_I = int(tm.sum()), int(im.sum())
            eng.alloc_outputs(n_th, n_I)
            tp = np.where(tm, np.cumsum(tm) - 1, -1)
            ip = np.where(im, np.cumsum(im) - 1, -1)
            eng.run(i0, n, tp, ip)
            total_ms += eng.stats()["step_ms"]
            if n_th and ex.get("vortex_planes"):
                # vortex configurations of the steps the caller asked for, computed where the phases are
                wanted = np.flatnonzero(ex["wanted"][i0:i1][tm])
                if wanted.size:
                    n_all = eng.vortex_configurations(int(wanted[0]), int(wanted[-1] - wanted[0] + 1))
                    dst0 = int(np.count_nonzero(ex["wanted"][:i0]))
                    sel = n_all if wanted[-1] - wanted[0] + 1 == wanted.size else n_all[wanted - wanted[0]]
                    if dst0 == 0 and sel.shape == ex["n_planes"].shape:
                        ex["n_planes"] = sel          # one shard, one chunk: the page-locked block itself, no second copy
                    else:
                        ex["n_planes"][dst0:dst0 + wanted.size, :, w0:w1] = sel
            if n_th and fetch_theta:
                first = th_idx[i0:i1][tm][0]
                eng.fetch_theta(0, n_th, out=th_host[lead + first: lead + first + n_th, :, w0:w1])
            if n_I:
                first = I_idx[i0:i1][im][0]
                eng.fetch_current(0, n_I, out=I_host[lead + first: lead + first + n_I, :, w0:w1])
            i0 = i1
        if ex.get("interval"):
            cnt, nsum, t_first, t_last = eng.observe_fetch()
            ex["count"] = cnt
            if w0 == 0 and w1 == ex["nsum"].shape[1]:
                # one shard holds every problem: hand the fetched (page-locked) arrays over as they are
                ex["nsum"], ex["theta_first"], ex["theta_latest"] = nsum, t_first, t_last
            else:
                ex["nsum"][:, w0:w1] = nsum
                ex["theta_first"][:, w0:w1] = t_first
                ex["theta_latest"][:, w0:w1] = t_last
            eng.observe_begin(0, 0)
        st = eng.stats()
        st["total_ms"] = total_ms
        st["problems"] = W
        stats_out[dev] = st
        ok = True
    finally:
        _release_engine(dev, key, eng, ok)


def _first_not_rank_one(specs, sh, i0, i1):
    for name, s in specs.items():
        if s.kind == RANK1 and not s.static:
            try:
                sh.amp(name, i0, i1)
            except NotRankOne:
                return name
    raise RuntimeError("no input failed the rank-one check on re-evaluation")


def _dense_for_device(name, table, tab):
    """(K, N, W) host values -> what the device expects for a DENSE input."""
    if name == "T":
        return np.sqrt(2.0 * table * tab.Rv[None, :, None])     # noise amplitude sqrt(2 T Rv)
    if name == "f":
        return table[:, tab.perm, :]
    return table


def device_time_evolution_core(problem, th_store_mask, I_store_mask, engine=None, shard=None, device=None,
                               initial_planes=True, noise_seed=None, extras=None):
    """
    Device replacement of time_evolution_core (reference: time_evolution.py:461-582), stencil width 3.
    Returns th_out, I_out of shape (Nj, W, n_stored + 2).
    initial_planes=False: the two leading planes (the initial conditions theta(-2), theta(-1) and their
    supercurrents) are left unwritten; time_evolution() asks for that when it is going to drop them unread
    (no voltage requested), which saves two passes over (Nj, W) arrays per plane on the host.
    shard=(w0, w1), device=d: integrate only problems [w0, w1) on GPU d and return (Nj, w1 - w0, .) arrays
    (used by distributed.compute_sharded, one process per GPU, which also passes the job-wide noise_seed).
    extras: dict describing what else the device should produce beside the planes; filled in place:
      interval, first       running observables (jj_observe_begin) -> count, nsum (Nf, W) int32, theta_first, theta_latest
      vortex_planes, wanted vortex configurations of the wanted stored steps -> n_planes (n_wanted, Nf, W) int32
      fetch_theta           False: the kept theta planes stay on the device (only the two initial planes are returned)
    """
    if getattr(problem, "stencil_width", 3) != 3:
        raise NotImplementedError("only stencil_width=3 is supported")
    circuit = problem.get_circuit()
    W = problem.get_problem_count()
    devices = getattr(problem, "devices", None)
    if devices is None:
        env = os.environ.get("JJ_DEVICES")
        devices = [int(d) for d in env.split(",")] if env else [0]
    # the dissection tree is made with one subtree per (SM, problem chunk) pair of the subdomain engine
    W_dev = (shard[1] - shard[0]) if shard is not None else -(-W // max(1, len(devices)))
    n_parts = None
    if os.environ.get("JJ_ENGINE", "auto") in ("auto", "subdomain") and not os.environ.get("JJ_SUBDOMAIN") \
            and engine in (None, _lib.JJ_ENGINE_AUTO, _lib.JJ_ENGINE_SUBDOMAIN):
        n_parts = subdomain_layout(circuit._Nf(), max(1, W_dev), _sm_count(devices[0] if device is None else device))[2]
    tab = _tables_for(circuit, problem._dt(), n_parts)
    Nj = tab.Nj
    th_mask = np.asarray(th_store_mask, dtype=bool)
    I_mask = np.asarray(I_store_mask, dtype=bool)
    specs = _classify_all(problem, tab)
    # every plane is written below: the two initial conditions here, the stored steps by the shards
    pin_dev = devices[0] if device is None else device
    ex = extras if extras is not None else {}
    if ex.get("interval"):
        single = len(devices) == 1 and shard is None
        ex["nsum"] = np.zeros((tab.Nf, W), dtype=np.int32)
        # (np.zeros pages are only touched when several shards fill their columns)
        ex["theta_first"], ex["theta_latest"], ex["count"] = (None, None, 0) if single else (np.zeros((Nj, W)), np.zeros((Nj, W)), 0)
    if ex.get("vortex_planes"):
        ex["wanted"] = np.asarray(ex.get("wanted", th_mask), dtype=bool) & th_mask
        ex["n_planes"] = np.zeros((int(ex["wanted"].sum()), tab.Nf, W), dtype=np.int32)
    keep_theta = th_mask.any() and ex.get("fetch_theta", True)
    # a caller that passes `extras` and does not want the initial planes gets arrays WITHOUT the two leading planes
    # (extras["lead"] = 0): 2 (Nj, W) planes less to pin and to allocate per output
    lead = 0 if (extras is not None and not initial_planes) else 2
    ex["lead"] = lead
    th_host = _pinned.empty((int(th_mask.sum()) + lead, Nj, W), pin_dev) if keep_theta else np.empty((lead, Nj, W))
    I_host = _pinned.empty((int(I_mask.sum()) + lead, Nj, W), pin_dev) if I_mask.any() else np.empty((lead, Nj, W))
    at_rest = getattr(problem, "starts_at_rest_with_zero_phases", None)
    at_rest = at_rest is not None and at_rest()          # no initial condition given: nothing is materialised for it
    if initial_planes:
        if at_rest:
            th_host[:2] = 0.0
        else:
            th_host[1] = problem.config_at_minus_1
            th_host[0] = problem.config_at_minus_2
    if not initial_planes:
        pass
    elif I_mask.any():
        # supercurrent of the initial conditions (reference: time_evolution.py:487); only read when currents
        # are stored or differentiated, so the two full-size sin passes are skipped otherwise
        if at_rest:
            I_host[:2] = problem._cp(np.zeros((Nj, 1)))
        else:
            I_host[1] = problem._cp(problem.config_at_minus_1)
            I_host[0] = problem._cp(problem.config_at_minus_2)
    else:
        I_host[:2] = 0.0
    if engine is None:
        engine = {"auto": _lib.JJ_ENGINE_AUTO, "streaming": _lib.JJ_ENGINE_STREAMING,
                  "subdomain": _lib.JJ_ENGINE_SUBDOMAIN}[os.environ.get("JJ_ENGINE", "auto")]
    bounds = shard_bounds(W, len(devices))
    stats = {}
    seed = resolve_noise_seed(problem, specs["T"].kind != ZERO) if noise_seed is None else int(noise_seed)
    jobs = [(dev, bounds[k], bounds[k + 1]) for k, dev in enumerate(devices) if bounds[k + 1] > bounds[k]]
    if shard is not None:
        jobs = [(devices[0] if device is None else device, shard[0], shard[1])] if shard[1] > shard[0] else []
    if len(jobs) == 0:
        pass
    elif len(jobs) == 1:
        dev, w0, w1 = jobs[0]
        _run_shard(problem, tab, specs, dev, w0, w1, th_mask, I_mask, th_host, I_host, engine, stats, seed, ex)
    else:
        errors = []

        def work(dev, w0, w1):
            try:
                _run_shard(problem, tab, specs, dev, w0, w1, th_mask, I_mask, th_host, I_host, engine, stats, seed, ex)
            except Exception as e:      # surfaced after join
                errors.append(e)
        threads = [threading.Thread(target=work, args=j) for j in jobs]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
    last_run_stats.clear()
    last_run_stats.update(stats)
    if shard is not None:
        th_host, I_host = th_host[:, :, shard[0]:shard[1]], I_host[:, :, shard[0]:shard[1]]
    # (plane, Nj, W) -> (Nj, W, plane) views, the reference's layout (quirk Q7)
    return np.moveaxis(th_host, 0, 2), np.moveaxis(I_host, 0, 2)


def _engine_kind(engine):
    if engine is None:
        engine = {"auto": _lib.JJ_ENGINE_AUTO, "streaming": _lib.JJ_ENGINE_STREAMING,
                  "subdomain": _lib.JJ_ENGINE_SUBDOMAIN}[os.environ.get("JJ_ENGINE", "auto")]
    return engine


def _n_parts_for(circuit, W_dev, device, engine):
    if os.environ.get("JJ_ENGINE", "auto") in ("auto", "subdomain") and not os.environ.get("JJ_SUBDOMAIN") \
            and engine in (None, _lib.JJ_ENGINE_AUTO, _lib.JJ_ENGINE_SUBDOMAIN):
        return subdomain_layout(circuit._Nf(), max(1, W_dev), _sm_count(device))[2]
    return None


def _anneal_shard(problem, tab, tab2, specs, dev, w0, w1, ann, adjust, out, engine_kind, stats_out):
    """Anneal problems [w0, w1) on one device (reference: time_evolution.py:1142-1191). The state, the factor and
    the stored theta planes stay on the device for the whole schedule; per interval the host receives one integer
    per problem (the vortex-mobility sum) and sends one temperature per problem."""
    W = w1 - w0
    dt, n_int, steps = ann["dt"], ann["interval_count"], ann["interval_steps"]
    seed, replay = ann["seed"], ann["noise_replay"]
    cpr = problem.current_phase_relation
    sh = _ShardInputs(specs, w0, w1)
    planes = np.arange(steps, dtype=np.int64)
    total_ms = 0.0
    launches = 0
    dense = any(s.kind == DENSE for s in specs.values())
    key, eng, run_kind = _engine_for(tab, cpr, dev, W, engine_kind, dense)
    ok = False
    try:
        eng.set_problem(W, dt, seed=seed, problem_offset=w0, engine=run_kind)     # theta(-1) = theta(-2) = 0
        _setup_sources(eng, specs, sh, tab)
        eng.set_source(_lib.JJ_SRC_T, _lib.JJ_KIND_RANK1, True, np.sqrt(2.0 * tab.Rv))
        eng.alloc_outputs(steps, 0)
        rule = ann.get("rule")
        if rule is not None and replay is None and os.environ.get("JJ_ANNEAL_HOST", "0") != "1":
            # the whole schedule in one call: amplitudes, restarts, mobility sums and the temperature rule stay on the
            # device (same arithmetic on the same integers as `adjust`, so the profiles are bit-identical)
            eng.upload_source(_lib.JJ_SRC_T, 0, np.zeros((1, W)))
            T_new, prof, ms = eng.anneal(0, n_int, steps, rule["upper"], rule["T_factor"], rule["norm"], out["T"][w0:w1])
            out["T"][w0:w1] = T_new
            out["profiles"][:, w0:w1] = prof
            total_ms += ms
            n_int = 0
        for i in range(n_int):
            T = out["T"][w0:w1]
            # the reference stops drawing noise once every temperature is numerically zero (time_evolution.py:512)
            amp = np.zeros(W) if np.allclose(T, 0) else np.sqrt(T)
            eng.upload_source(_lib.JJ_SRC_T, 0, amp[None, :])
            if replay is not None:
                Z = replay(i) if callable(replay) else replay[i]
                eng.upload_noise(i * steps, np.asarray(Z)[:, :, w0:w1])
            if i > 0:
                eng.restart_at_rest()
            eng.run(i * steps, steps, planes, None)
            total_ms += eng.stats()["step_ms"]
            sums = eng.vortex_mobility_sums(0, steps)
            out["T"][w0:w1] = adjust(sums, i, T)
            out["profiles"][i, w0:w1] = out["T"][w0:w1]
        launches += eng.stats()["kernel_launches"]
        # closing runs at T = 0 with half the time step (reference: time_evolution.py:1176-1183): another factor, i.e.
        # another engine; the phases go over device to device
        key2, eng2, run_kind2 = _engine_for(tab2, cpr, dev, W, engine_kind, dense)
        try:
            eng2.set_problem(W, dt / 2, seed=seed, problem_offset=w0, engine=run_kind2)
            eng2.adopt_state_at_rest(eng)
        except BaseException:
            _release_engine(dev, key2, eng2, False)
            raise
        ok = True
    finally:
        _release_engine(dev, key, eng, ok)
    key, eng, run_kind = key2, eng2, run_kind2
    ok = False
    try:
        _setup_sources(eng, specs, sh, tab2)
        eng.set_source(_lib.JJ_SRC_T, _lib.JJ_KIND_ZERO, True)
        eng.alloc_outputs(0, 0)
        for r in range(ann["final_runs"]):
            if r > 0:
                eng.restart_at_rest()
            eng.run(r * steps, steps, None, None)
            total_ms += eng.stats()["step_ms"]
        whole = w0 == 0 and w1 == out["theta"].shape[1]          # one shard: the phases land in the result array itself
        th, _ = eng.get_state(previous=False, out=out["theta"] if whole else None)
        if th is not out["theta"]:
            out["theta"][:, w0:w1] = th
        out["n"][:, w0:w1] = eng.vortex_configuration(-1)
        st = eng.stats()
        st["total_ms"] = total_ms
        st["problems"] = W
        st["kernel_launches"] += launches
        stats_out[dev] = st
        ok = True
    finally:
        _release_engine(dev, key, eng, ok)


def device_annealing(problem, T0, adjust, interval_count, final_runs=5, engine=None, rule=None):
    """
    Device-resident annealing schedule around the stepping loop (reference: time_evolution.py:1142-1191).

    problem : the TimeEvolutionProblem the reference's loop would re-run (circuit, dt, interval_steps steps, flux,
              current sources; its temperature is ignored). T0 : (W,) start temperatures.
    adjust(sums, i, T) -> new T : the temperature rule, given the exact integer mobility sums of interval i.
    rule : the same rule as data, dict(upper (interval_count,), T_factor, norm): new T = T / T_factor where
           sums / norm > upper[i], else T * T_factor. When given (and no noise is replayed) the schedule runs in one
           device call per shard (jj_anneal) instead of one host round trip per interval.
    Returns dict(profiles (interval_count, W), theta (Nj, W), n (Nf, W), stats).
    """
    if getattr(problem, "stencil_width", 3) != 3:
        raise NotImplementedError("only stencil_width=3 is supported")
    circuit = problem.get_circuit()
    W = problem.get_problem_count()
    devices = getattr(problem, "devices", None)
    if devices is None:
        env = os.environ.get("JJ_DEVICES")
        devices = [int(d) for d in env.split(",")] if env else [0]
    engine = _engine_kind(engine)
    n_parts = _n_parts_for(circuit, -(-W // max(1, len(devices))), devices[0], engine)
    dt = problem._dt()
    tab = _tables_for(circuit, dt, n_parts)
    tab2 = _tables_for(circuit, dt / 2, n_parts)
    specs = _classify_all(problem, tab)
    specs.pop("T")
    if specs["Vs"].kind == DENSE:
        raise NotImplementedError("annealing with dense voltage sources is not supported")
    out = dict(T=np.array(T0, dtype=np.double).reshape(W).copy(), profiles=np.zeros((interval_count, W)),
               theta=_pinned.empty((tab.Nj, W), devices[0]), n=np.zeros((tab.Nf, W), dtype=int))
    ann = dict(dt=dt, interval_count=int(interval_count), interval_steps=problem._Nt(), final_runs=int(final_runs),
               seed=resolve_noise_seed(problem),
               noise_replay=getattr(problem, "noise_replay", None), rule=rule)
    bounds = shard_bounds(W, len(devices))
    jobs = [(dev, bounds[k], bounds[k + 1]) for k, dev in enumerate(devices) if bounds[k + 1] > bounds[k]]
    stats, errors = {}, []

    def work(dev, w0, w1):
        try:
            _anneal_shard(problem, tab, tab2, specs, dev, w0, w1, ann, adjust, out, engine, stats)
        except Exception as e:      # surfaced after join
            errors.append(e)
    if len(jobs) == 1:
        _anneal_shard(problem, tab, tab2, specs, *jobs[0], ann, adjust, out, engine, stats)
    else:
        threads = [threading.Thread(target=work, args=j) for j in jobs]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
    last_run_stats.clear()
    last_run_stats.update(stats)
    out["stats"] = stats
    return out


_sm_cache = {}


def _sm_count(device):
    """Multiprocessor count of a device (through the C ABI; 148 = B200 if the library cannot tell yet)."""
    if device not in _sm_cache:
        try:
            _sm_cache[device] = int(_lib.load().jj_sm_count(int(device)))
        except Exception:
            _sm_cache[device] = 0
    return _sm_cache[device] if _sm_cache[device] > 0 else 148


def shard_bounds(W, n_shards):
    """Contiguous shards of the problem axis whose starts are multiples of 4 (Philox groups)."""
    groups = (W + 3) // 4
    per = [(groups // n_shards + (1 if k < groups % n_shards else 0)) for k in range(n_shards)]
    b = np.minimum(np.concatenate(([0], np.cumsum(per))) * 4, W)
    return [int(v) for v in b]
