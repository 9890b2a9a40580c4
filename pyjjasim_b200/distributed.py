"""
Multi-GPU use of the time-evolution path: shard the problem axis W, one process per GPU.

Problems never interact (every operation of the stepping loop is column-wise in W,
reference: time_evolution.py:523-580), so the batch is cut into contiguous shards, each rank integrates its
shard on its own GPU with its own replica of the factor, and there is NO collective on the step path. The
only communication is one gather of the stored planes at the end (NCCL all_gather on GPU boxes, gloo in CPU
tests). Shard starts are multiples of 4 so the counter-based noise, keyed by the global problem index, is
identical however the batch is sharded.

Single-process alternative: ``TimeEvolutionProblem(..., devices=[0, 1, ...])`` drives several GPUs from host
threads (engine.device_time_evolution_core).
"""
import time

import numpy as np

from .engine import shard_bounds

__all__ = ["shard_for_rank", "gather_problem_axis", "compute_sharded", "last_gather_seconds"]

last_gather_seconds = None      # wall time of the final gather of the last compute_sharded call on this rank


def shard_for_rank(W, rank, world_size):
    """[w0, w1) of this rank."""
    b = shard_bounds(W, world_size)
    return b[rank], b[rank + 1]


def gather_problem_axis(local, W, group=None):
    """
    All-gather arrays sharded along axis 1 (the problem axis): ``local`` is (N, w1 - w0, K) on each rank;
    returns the full (N, W, K) array on every rank. Uses torch.distributed (any backend).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    b = shard_bounds(W, world)
    widths = [b[r + 1] - b[r] for r in range(world)]
    wmax = max(widths)
    N, _, K = local.shape
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((wmax, N, K), dtype=torch.float64, device=dev)
    buf[: local.shape[1]] = torch.from_numpy(np.ascontiguousarray(np.moveaxis(local, 1, 0))).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = np.empty((N, W, K))
    for r in range(world):
        if widths[r]:
            out[:, b[r]:b[r + 1], :] = np.moveaxis(parts[r][: widths[r]].cpu().numpy(), 0, 1)
    return out


def compute_sharded(problem, device=None, group=None, core=None):
    """
    SPMD time evolution: every rank calls this with the same TimeEvolutionProblem; each integrates its
    shard of the problem axis on ``device`` (default: LOCAL_RANK) and all ranks return the complete
    TimeEvolutionResult. ``core`` replaces the device core (tests inject the CPU oracle here).
    """
    import os
    import torch.distributed as dist
    from .time_evolution import time_evolution as _time_evolution
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    W = problem.get_problem_count()
    w0, w1 = shard_for_rank(W, rank, world)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))

    def sharded_core(prob, th_mask, I_mask, extras=None):
        if core is not None:
            th, I = core(prob, th_mask, I_mask, w0, w1)
        else:
            from .engine import device_time_evolution_core, resolve_noise_seed
            # one Philox seed for the whole job: rank 0 resolves it (explicit noise_seed, or a fresh draw), all use it
            box = [resolve_noise_seed(prob) if rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)
            th, I = device_time_evolution_core(prob, th_mask, I_mask, shard=(w0, w1), device=device, noise_seed=box[0],
                                               initial_planes=True if extras is None else bool(extras.get("initial_planes", True)),
                                               extras=extras)
        global last_gather_seconds
        t0 = time.perf_counter()
        out = gather_problem_axis(th, W, group), gather_problem_axis(I, W, group)
        if extras is not None:
            # what the device produced beside the planes, gathered along the problem axis as well (every rank filled
            # the columns of its shard): vortex sums and phase marks of the running observables, device-side vortex planes
            for key in ("nsum", "theta_first", "theta_latest"):
                if key in extras:
                    full = gather_problem_axis(np.asarray(extras[key][:, w0:w1, None], dtype=np.double), W, group)[:, :, 0]
                    extras[key] = full.astype(extras[key].dtype)
            if extras.get("n_planes") is not None:
                loc = np.moveaxis(extras["n_planes"][:, :, w0:w1], 0, 2).astype(np.double)        # (Nf, w, K)
                extras["n_planes"] = np.moveaxis(gather_problem_axis(loc, W, group), 2, 0).astype(np.int32)
        last_gather_seconds = time.perf_counter() - t0
        return out

    sharded_core.accepts_extras = core is None
    return _time_evolution(problem, core=sharded_core)
