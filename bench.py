#!/usr/bin/env python
"""
Benchmark of the time-evolution hot path (BASELINE.json metric: junction-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at N=1 (BASELINE.json configs[1], SURVEY.md section 8d "cfg2"): SquareArray(100,100), f = 0.1,
thermal noise with a 256-temperature batch T = geomspace(1e-2, 1, 256), dt = 0.5, Philox seed 1234.
One bench "step" = INNER consecutive time steps of the whole batch (one jj_run call, state resident in HBM).
N > 1 (torchrun, one rank per GPU): every rank integrates its own 256-temperature shard of a 256*N batch
(weak scaling, no collective on the step path); times are device times, max over ranks.

Printed JSON keys follow the driver contract; `value` = device-resident throughput, `e2e` = the same
metric through TimeEvolutionProblem.compute() with host buffers (H2D of the initial phases and source
tables and D2H of the stored phases inside the timed region).
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
DT, FRUST, SEED = 0.5, 0.1, 1234


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


def workload(rank, world):
    import pyjjasim_b200 as pj
    a = pj.SquareArray(NX, NX)
    Wtot = W_PER_GPU * world
    T_all = np.geomspace(1e-2, 1.0, Wtot)
    w0 = rank * W_PER_GPU
    return a, T_all[w0:w0 + W_PER_GPU], w0


def algorithmic_bytes_per_time_step(tab, W):
    """8 W (4 Nj + 2 Nf) + 32 W Nf + 24 nnz(L)."""
    nnzL = tab.factor.nnz_L
    return 8 * W * (4 * tab.Nj + 2 * tab.Nf) + 32 * W * tab.Nf + 24 * nnzL


def run_ours(args, rank, world, local_rank, dist):
    import torch
    import pyjjasim_b200 as pj
    from pyjjasim_b200 import engine, _lib
    torch.cuda.set_device(local_rank)
    a, T, w0 = workload(rank, world)
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
    total_ms = float(np.sum(dev_ms))
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
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
            prob = pj.TimeEvolutionProblem(a, time_step=DT, time_step_count=INNER, external_flux=FRUST,
                                           temperature=T[None, :, None], store_time_steps=[INNER // 3, INNER - 1],
                                           store_current=False, store_voltage=False, config_at_minus_1=th0,
                                           noise_seed=SEED)
            if os.environ.get("JJ_BENCH_E2E_PROFILE") and r == reps + e2e_warm - 1:   # where the host time of the last repeat goes
                import cProfile, pstats
                pr = cProfile.Profile(); pr.enable(); res = prob.compute(); pr.disable()
                pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(14)
            else:
                res = prob.compute()
            e2e_t.append(time.perf_counter() - t1)
            th0 = np.ascontiguousarray(res.theta[:, :, -1])
        e2e_s = float(np.mean(e2e_t[e2e_warm:]))
        if dist is not None:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": tab.Nj * W * world * INNER / e2e_s, "unit": "junction-steps/s",
               "h2d_bytes_per_step": int(2 * tab.Nj * W * 8 + 2 * W * 8), "d2h_bytes_per_step": int(2 * tab.Nj * W * 8),
               "seconds_per_step": e2e_s}
    except Exception as e:
        e2e = {"error": repr(e)}

    out = {"metric": "junction-steps/sec (junctions x timesteps x problems)", "value": value,
           "unit": "junction-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"cfg2: SquareArray({NX},{NX}) f=0.1 thermal noise, {W} temperatures per GPU, dt=0.5",
                      "Nj": tab.Nj, "Nf": tab.Nf, "problems_per_gpu": W, "time_steps_per_step": INNER,
                      "l2": "256 MiB buffer written between timed steps (L2 flush)", "noise": "device Philox4x32-10, seed 1234",
                      "engine": {1: "streaming", 3: "subdomain"}.get(st["engine"], str(st["engine"])),
                      "subdomains_or_cluster": st["cluster_size"], "problems_per_block": st["tile_problems"]},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "peak_source": peak_src,
                        "bytes_model": "SURVEY 8(d): (8W(4Nj+2Nf) + 32 W Nf + 24 nnz(L)) x time steps per launch"},
           "wall_s": wall}
    return out


def cpu_reference(steps, warmup, W_cpu=None, nt=None):
    """The CPU path (numpy/scipy oracle port of the reference's loop) on the host cores, bounded sample."""
    import warnings
    import pyjjasim_b200 as pj
    from oracle import oracle
    a = pj.SquareArray(NX, NX)
    W = W_cpu or int(os.environ.get("JJ_BENCH_CPU_W", "64"))
    nt = nt or int(os.environ.get("JJ_BENCH_CPU_NT", "12"))
    T = np.geomspace(1e-2, 1.0, W)[None, :, None]
    A, L = a.get_cycle_matrix(), a._L()
    rng = np.random.RandomState(SEED)
    th = np.zeros((a._Nj(), W))
    times = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # the factorisation is a one-off setup cost: time it separately and subtract it
        import scipy.sparse
        import scipy.sparse.linalg
        t0 = time.perf_counter()
        Rv, Cv = 1 / (DT * a._R()), a._C() / DT ** 2
        scipy.sparse.linalg.factorized(A @ (L + scipy.sparse.diags(1.0 / (Cv + Rv), 0)) @ A.T)
        t_setup = time.perf_counter() - t0
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            th_out, _ = oracle.time_evolution_core(A, a._Ic(), a._R(), a._C(), L, DT, nt, W, f=FRUST, T=T,
                                                   theta_m1=th, th_store_mask=np.arange(nt) == nt - 1,
                                                   I_store_mask=np.zeros(nt, bool), rng=rng)
            dt_run = time.perf_counter() - t0 - t_setup
            th = th_out[:, :, -1]
            if k >= warmup:
                times.append(max(dt_run, 1e-9))
    js = a._Nj() * W * nt
    return {"value": js / float(np.mean(times)), "unit": "junction-steps/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"SquareArray({NX},{NX}) f=0.1 T>0, {W} of 256 problems x {nt} time steps per step, "
                      f"{steps} steps, factorisation ({t_setup:.2f} s) subtracted; numpy/scipy(SuperLU) as the reference uses",
            "seconds_per_step": float(np.mean(times)), "threads": "numpy single thread + OpenBLAS default inside SuperLU"}


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
        cb = cpu_reference(max(1, min(args.steps, 3)), min(args.warmup, 1))
        out = {"impl": "reference", "metric": "junction-steps/sec (junctions x timesteps x problems)",
               "value": cb["value"], "unit": "junction-steps/s", "n_gpus": args.gpus, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"cfg2: SquareArray({NX},{NX}) f=0.1 thermal noise, dt=0.5 (CPU sample: {cb['sample']})"},
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
            out["cpu_baseline"] = {"skipped": True} if os.environ.get("JJ_BENCH_SKIP_E2E") else cpu_reference(1, 0)
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
