#!/usr/bin/env python
"""
Benchmark of the time-evolution hot path (BASELINE.json metric: junction-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (BASELINE.json configs[1], SURVEY.md section 8d "cfg2"): SquareArray(100,100), f = 0.1, thermal
noise with a 256-temperature batch T = geomspace(1e-2, 1, 256), dt = 0.5, Philox seed 1234. One bench "step" = INNER
consecutive time steps of the whole batch (one jj_run call, state resident in HBM). N > 1 (torchrun, one rank per
GPU): every rank integrates its own 256-temperature shard of a 256*N batch (weak scaling, no collective on the step
path); times are device times, max over ranks.

Printed JSON keys follow the driver contract; `value` = device-resident throughput, `e2e` = the same metric through
TimeEvolutionProblem.compute() with host buffers (H2D of the initial phases and source tables and D2H of the stored
phases inside the timed region). Beyond the contract the line carries
  per_config  the other BASELINE.json configurations measured in the same run (device value, e2e through the public
              API, roofline fraction, engine): cfg1, cfg3 (512 problems per GPU), cfg4 (512 on one GPU; the 4096-problem
              sweep cut into 4096/N shards at N > 1), cfg5 (64 problems cut into 64/N shards);
  shard_check at N > 1: every rank re-computes the first problems of its neighbour's shard as a separate small batch
              with the same global problem indices and the ranks compare (the noise is keyed by the global index, so
              any sharding gives the same trajectories), and rank 0 times the final gather of compute_sharded.

--impl reference: the reference's CPU path (numpy/scipy port in oracle/, bitwise equal to the reference) on the host
cores, on the same workload with the same problems per shard; a bench step is a bounded sample of CPU_NT time steps,
timed whole (factorisation included, as in every compute() of the reference).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INNER = int(os.environ.get("JJ_BENCH_INNER", "1000"))      # time steps per bench step
NX = int(os.environ.get("JJ_BENCH_NX", "100"))
W_PER_GPU = int(os.environ.get("JJ_BENCH_W", "256"))
CPU_NT = int(os.environ.get("JJ_BENCH_CPU_NT", "4"))       # time steps of one bench step of the CPU arm
DT, FRUST, SEED = 0.5, 0.1, 1234
METRIC = "junction-steps/sec (junctions x timesteps x problems)"


def workload_name(W):
    return f"cfg2: SquareArray({NX},{NX}) f=0.1 thermal noise, {W} temperatures per GPU, dt=0.5"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:      # NVML missing: report that instead of inventing numbers
            self.reasons.add("nvml_unavailable:" + type(e).__name__)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons)}


def cfg2_problem(pj, a, T, Nt, th0=None, store=None):
    return pj.TimeEvolutionProblem(a, time_step=DT, time_step_count=Nt, external_flux=FRUST,
                                   temperature=T[None, :, None], store_time_steps=store if store is not None else [Nt // 3, Nt - 1],
                                   store_current=False, store_voltage=False, config_at_minus_1=th0, noise_seed=SEED)


def algorithmic_bytes_per_time_step(tab, W):
    """8 W (4 Nj + 2 Nf) + 32 W Nf + 24 nnz(L)."""
    return 8 * W * (4 * tab.Nj + 2 * tab.Nf) + 32 * W * tab.Nf + 24 * tab.factor.nnz_L


def max_over_ranks(x, dist):
    if dist is None:
        return float(x)
    import torch
    t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_ours(args, rank, world, local_rank, dist):
    import torch
    import pyjjasim_b200 as pj
    from pyjjasim_b200 import engine, _lib
    torch.cuda.set_device(local_rank)
    a = pj.SquareArray(NX, NX)
    T_all = np.geomspace(1e-2, 1.0, W_PER_GPU * world)
    w0 = rank * W_PER_GPU
    T = T_all[w0:w0 + W_PER_GPU]
    W = T.size
    n_parts = None
    if os.environ.get("JJ_ENGINE", "auto") in ("auto", "subdomain") and not os.environ.get("JJ_SUBDOMAIN"):
        n_parts = engine.subdomain_layout(a._Nf(), W, engine._sm_count(local_rank))[2]
    if os.environ.get("JJ_BENCH_NPARTS"):            # experiments: another decomposition / problem count
        n_parts = int(os.environ["JJ_BENCH_NPARTS"])
    tab = engine.CircuitTables(a, DT, n_parts=n_parts)
    eng = engine.DeviceEngine(local_rank)
    kind = {"auto": _lib.JJ_ENGINE_AUTO, "streaming": _lib.JJ_ENGINE_STREAMING,
            "subdomain": _lib.JJ_ENGINE_SUBDOMAIN}[os.environ.get("JJ_ENGINE", "auto")]
    cfg = tab.choose_subdomain(W) if kind in (_lib.JJ_ENGINE_AUTO, _lib.JJ_ENGINE_SUBDOMAIN) else None
    eng.set_circuit(tab, pj.DefaultCPR(), with_program=cfg is None)
    if cfg is not None:
        eng.set_subdomain(*cfg)
        kind = _lib.JJ_ENGINE_SUBDOMAIN
    eng.set_problem(W, DT, seed=SEED, problem_offset=w0, engine=kind)
    eng.set_source(_lib.JJ_SRC_F, _lib.JJ_KIND_RANK1, True, np.ones(tab.Nf))
    eng.upload_source(_lib.JJ_SRC_F, 0, np.full((1, W), FRUST))
    if os.environ.get("JJ_BENCH_T0"):          # experiment only: the same workload without thermal noise
        eng.set_source(_lib.JJ_SRC_T, _lib.JJ_KIND_ZERO, True)
    else:
        eng.set_source(_lib.JJ_SRC_T, _lib.JJ_KIND_RANK1, True, np.sqrt(2.0 * np.ones(tab.Nj) * tab.Rv))
        eng.upload_source(_lib.JJ_SRC_T, 0, np.sqrt(T)[None, :])
    eng.alloc_outputs(1, 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    step_no = [0]

    def one_step():
        plane = -np.ones(INNER, dtype=np.int64)
        plane[-1] = 0                                     # keep the last phases of the step (device resident)
        eng.run(step_no[0] * INNER, INNER, plane, None)
        step_no[0] += 1
        return eng.stats()["step_ms"]

    for _ in range(args.warmup):
        one_step()
    launches0 = eng.stats()["kernel_launches"]
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)                                    # L2 flush between timed iterations (not timed: events bracket jj_run)
        dev_ms.append(one_step())
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.result()
    st = eng.stats()
    launches = st["kernel_launches"] - launches0
    total_ms = max_over_ranks(np.sum(dev_ms), dist)
    th_last = eng.fetch_theta(0, 1)
    assert np.all(np.isfinite(th_last))
    js = tab.Nj * W * world * INNER * args.steps
    value = js / (total_ms * 1e-3)
    # roofline of the dominant kernel(s): algorithmic bytes of one jj_run / its device time
    peak, peak_src = measured_peak()
    bytes_run = algorithmic_bytes_per_time_step(tab, W) * INNER
    achieved = bytes_run / (np.mean(dev_ms) * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp):
        try:
            traffic = json.load(open(tp)).get(str(st["engine"]))
            if isinstance(traffic, dict):      # measured DRAM bytes per time step (ncu) x time steps per launch
                traffic = traffic["bytes_per_time_step"] * INNER
        except Exception:
            traffic = None
    eng.close()

    # end to end through the public API with host buffers
    e2e = None
    try:
        if os.environ.get("JJ_BENCH_SKIP_E2E"):
            raise RuntimeError("skipped (JJ_BENCH_SKIP_E2E)")
        os.environ["JJ_DEVICES"] = str(local_rank)
        th0 = np.zeros((tab.Nj, W))
        # warm-up calls as for the device leg: the first pays the one-off factorisation, the first two pin their result
        # blocks (the pool hands them out again once the previous result is dropped)
        reps, e2e_warm, e2e_t = max(1, min(3, args.steps)), max(2, min(3, args.warmup)), []
        for r in range(reps + e2e_warm):
            t1 = time.perf_counter()
            res = cfg2_problem(pj, a, T, INNER, th0).compute()
            e2e_t.append(time.perf_counter() - t1)
            th0 = np.ascontiguousarray(res.theta[:, :, -1])
        e2e_s = max_over_ranks(np.mean(e2e_t[e2e_warm:]), dist)
        e2e = {"value": tab.Nj * W * world * INNER / e2e_s, "unit": "junction-steps/s",
               "h2d_bytes_per_step": int(2 * tab.Nj * W * 8 + 2 * W * 8), "d2h_bytes_per_step": int(2 * tab.Nj * W * 8),
               "seconds_per_step": e2e_s}
    except Exception as e:
        e2e = {"error": repr(e)}

    out = {"metric": METRIC, "value": value,
           "unit": "junction-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(W),
                      "Nj": tab.Nj, "Nf": tab.Nf, "problems_per_gpu": W, "time_steps_per_step": INNER,
                      "l2": "256 MiB buffer written between timed steps (L2 flush)", "noise": "device Philox4x32-10, seed 1234",
                      "engine": {1: "streaming", 3: "subdomain"}.get(st["engine"], str(st["engine"])),
                      "subdomains_or_cluster": st["cluster_size"], "problems_per_block": st["tile_problems"]},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "peak_source": peak_src,
                        "bytes_model": "SURVEY 8(d): (8W(4Nj+2Nf) + 32 W Nf + 24 nnz(L)) x time steps per launch"},
           "wall_s": wall}
    if not os.environ.get("JJ_BENCH_SKIP_CONFIGS"):
        out["per_config"] = per_config(rank, world, local_rank, dist, peak)
    if dist is not None and not os.environ.get("JJ_BENCH_SKIP_SHARD_CHECK"):
        try:
            out["shard_check"] = shard_check(pj, a, T_all, rank, world, local_rank, dist)
        except Exception as e:
            out["shard_check"] = {"error": repr(e)}
    return out


# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations, through the public API
# ------------------------------------------------------------------------------------------------
def named_config(pj, name, rank, world):
    """-> (circuit, constructor keywords, time steps, this rank's problems [w0, w1) of W_total, note)"""
    if name == "cfg1":
        a = pj.SquareArray(20, 20)
        W = 32
        Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0, 2, W))
        return a, dict(time_step=0.05, current_sources=Is), 10000, (0, W if rank == 0 else 0), W, "32 problems on one GPU (does not shard)"
    if name == "cfg3":
        a = pj.HoneycombArray(200, 200)
        W = 512 * world
        w0 = 512 * rank
        Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0.1, 1.5, W)[w0:w0 + 512])
        return a, dict(time_step=0.05, external_flux=0.1, current_sources=Is), 100, (w0, w0 + 512), W, "512 problems per GPU"
    if name == "cfg4":
        a = pj.SquareArray(256, 256)
        a.set_capacitance(1.0)
        W = 512 if world == 1 else 4096
        per = W // world
        w0 = per * rank
        IDC = np.repeat(np.linspace(0, 2, 64), 64)[:W] if W == 4096 else np.linspace(0, 2, W)
        IA = np.tile(np.linspace(0, 3, 64), 64)[:W] if W == 4096 else np.linspace(0, 3, W)
        idc, ia = IDC[w0:w0 + per], IA[w0:w0 + per]
        Is = pj.RankOneSource(a.current_base(angle=0), lambda i: idc + ia * np.sin(0.25 * i * 0.05), problem_count=per)
        note = "one GPU's share (512) of the 4096-problem sweep" if world == 1 else f"4096 problems in shards of {per}"
        return a, dict(time_step=0.05, current_sources=Is, temperature=0.01 * np.ones((1, per, 1)), noise_seed=SEED), 100, (w0, w0 + per), W, note
    if name == "cfg5":
        a = pj.SquareArray(1000, 1000)
        a.set_inductance(1.0)
        a.set_capacitance(1.0)
        W = 64
        per = W // world
        w0 = per * rank
        Is = pj.RankOneSource(a.current_base(angle=0), np.linspace(0.5, 1.5, W)[w0:w0 + per])
        return a, dict(time_step=0.05, external_flux=0.05, current_sources=Is), 20, (w0, w0 + per), W, f"64 problems in shards of {per}"
    raise KeyError(name)


def per_config(rank, world, local_rank, dist, peak):
    import pyjjasim_b200 as pj
    from pyjjasim_b200 import engine
    os.environ["JJ_DEVICES"] = str(local_rank)
    names = os.environ.get("JJ_BENCH_CONFIGS", "cfg1,cfg3,cfg4,cfg5").split(",")
    out = {}
    for name in names:
        # the timings are collected locally first; the collectives below are reached by every rank whatever happened
        rec, dev_s, wall, js, err = {}, 0.0, 0.0, 0.0, None
        try:
            a, kw, Nt, (w0, w1), W_total, note = named_config(pj, name, rank, world)
            rec = {"problems_total": W_total, "problems_this_gpu": w1 - w0, "time_steps": Nt, "note": note,
                   "Nj": a._Nj(), "Nf": a._Nf(), "unit": "junction-steps/s", "roofline_frac": None}
            js = float(a._Nj()) * W_total * Nt
            if w1 > w0:
                t0 = time.perf_counter()
                prob = pj.TimeEvolutionProblem(a, time_step_count=Nt, store_time_steps=[Nt - 1], store_current=False,
                                               store_voltage=False, **kw)
                res = prob.compute()               # first call: ordering, factorisation, plans, upload
                t1 = time.perf_counter()
                del res
                res = prob.compute()               # warm: result block pinned and pooled, engine cached
                del res
                t1b = time.perf_counter()
                res = prob.compute()               # the timed call (steady state of repeated calls, like the cfg2 e2e leg)
                t2 = time.perf_counter()
                st = list(engine.last_run_stats.values())[0]
                assert np.all(np.isfinite(res.theta))
                dev_s, wall = st["total_ms"] * 1e-3, t2 - t1b
                tab = engine._tables_for(a, kw["time_step"], engine._n_parts_for(a, w1 - w0, local_rank, None))
                rec.update(engine={1: "streaming", 3: "subdomain"}.get(st["engine"]), subdomains=st["cluster_size"],
                           setup_s=round((t1 - t0) - (t2 - t1b), 2), device_us_per_time_step=round(dev_s * 1e6 / Nt, 1),
                           roofline_frac=algorithmic_bytes_per_time_step(tab, w1 - w0) * Nt / dev_s / 1e9 / peak)
                del res, prob
        except Exception as e:
            err = repr(e)
        finally:
            engine._tables_cache.clear()
        if name != "cfg1":
            dev_s, wall = max_over_ranks(dev_s, dist), max_over_ranks(wall, dist)
            err_any = max_over_ranks(1.0 if err else 0.0, dist) > 0
        else:
            err_any = err is not None
            if rank != 0:
                continue
        if err_any or dev_s <= 0:
            out[name] = {"error": err or "failed on another rank"}
        else:
            rec.update(value=js / dev_s, e2e=js / wall)
            out[name] = rec
    return out


def shard_check(pj, a, T_all, rank, world, local_rank, dist):
    """Sharding must not change results: the noise is keyed by the GLOBAL problem index. Every rank integrates (a) its own
    shard and (b) the first 8 problems of the next rank's shard as a batch of their own with the same global indices;
    the ranks exchange (b) and compare it with their (a). (The dissection differs with the batch size, so the comparison
    is to round-off over a short horizon, not bitwise.) Rank 0 also times the final gather of compute_sharded."""
    import torch
    from pyjjasim_b200 import engine, distributed
    Nt, nb = 20, 8
    os.environ["JJ_DEVICES"] = str(local_rank)
    Wtot = T_all.size

    def run(w0, w1):
        prob = cfg2_problem(pj, a, T_all, Nt, store=[Nt - 1])
        th, _ = engine.device_time_evolution_core(prob, prob.store_time_steps, np.zeros(Nt, bool), shard=(w0, w1),
                                                  device=local_rank, initial_planes=False, noise_seed=SEED)
        return np.ascontiguousarray(th[:, :, -1])
    b = engine.shard_bounds(Wtot, world)
    own = run(b[rank], b[rank + 1])
    nxt = (rank + 1) % world
    theirs = run(b[nxt], b[nxt] + nb)                       # problems of the next rank, computed here as a small batch
    buf = torch.from_numpy(theirs).cuda()
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    mine_by_prev = parts[(rank - 1) % world].cpu().numpy()  # my first problems, as computed by the previous rank
    diff = float(np.max(np.abs(mine_by_prev - own[:, :nb])))
    t = torch.tensor([diff], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    diff = float(t.item())
    # final gather of the public multi-process driver
    prob = cfg2_problem(pj, a, T_all, Nt, store=[Nt // 2, Nt - 1])
    res = distributed.compute_sharded(prob, device=local_rank)
    full_ok = bool(np.max(np.abs(res.theta[:, b[rank]:b[rank + 1], -1] - own)) <= 1e-9)
    ok = torch.tensor([1.0 if (full_ok and diff <= 1e-9) else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return {"ok": bool(ok.item() > 0), "max_abs_dtheta_across_shardings": diff, "tolerance": 1e-9, "time_steps": Nt,
            "problems_compared_per_rank": nb,
            "gather_seconds": distributed.last_gather_seconds, "gather_bytes_per_rank": int(res.theta.nbytes)}


# ------------------------------------------------------------------------------------------------
# the CPU arm
# ------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, W=None, nt=None):
    """The reference's CPU path (numpy/scipy oracle port of its loop, bitwise equal to it) on the host cores: `steps`
    timed calls after `warmup` untimed ones, each the whole core (assembly, factorisation, nt time steps) on W of the
    workload's temperatures, state carried over from call to call."""
    import warnings
    import pyjjasim_b200 as pj
    from oracle import oracle
    a = pj.SquareArray(NX, NX)
    W = W or W_PER_GPU
    nt = nt or CPU_NT
    T = np.geomspace(1e-2, 1.0, W)[None, :, None]
    A, L = a.get_cycle_matrix(), a._L()
    rng = np.random.RandomState(SEED)
    th = np.zeros((a._Nj(), W))
    times = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            th_out, _ = oracle.time_evolution_core(A, a._Ic(), a._R(), a._C(), L, DT, nt, W, f=FRUST, T=T,
                                                   theta_m1=th, th_store_mask=np.arange(nt) == nt - 1,
                                                   I_store_mask=np.zeros(nt, bool), rng=rng)
            dt_run = time.perf_counter() - t0
            th = th_out[:, :, -1]
            if k >= warmup:
                times.append(dt_run)
    js = a._Nj() * W * nt
    return {"value": js / float(np.mean(times)), "unit": "junction-steps/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"SquareArray({NX},{NX}) f=0.1 T>0, all {W} problems of a shard x {nt} time steps per step, "
                      f"{steps} timed steps after {warmup} warm-up steps, each a whole core call (assembly + "
                      "factorisation + time steps); numpy/scipy(SuperLU) as the reference uses",
            "seconds_per_step": float(np.mean(times)), "time_steps_per_step": nt,
            "threads": "numpy single thread + OpenBLAS default inside SuperLU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference(args.steps, args.warmup)
        out = {"impl": "reference", "metric": METRIC,
               "value": cb["value"], "unit": "junction-steps/s", "n_gpus": args.gpus, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(W_PER_GPU), "Nj": 2 * NX * (NX - 1), "Nf": (NX - 1) ** 2,
                          "problems_per_gpu": W_PER_GPU, "time_steps_per_step": cb["time_steps_per_step"],
                          "noise": "numpy RandomState(1234), the reference's draw sequence"},
               "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": "junction-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return

    dist = None
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner at communicator
    # creation) is sent to stderr by pointing fd 1 at fd 2 until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local_rank)
        # NCCL's own banner ("NCCL version ...") goes to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = td
    out = run_ours(args, rank, world, local_rank, dist)
    if rank == 0:
        try:
            out["cpu_baseline"] = {"skipped": True} if os.environ.get("JJ_BENCH_SKIP_E2E") else cpu_reference(3, 1)
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
