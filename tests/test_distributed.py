"""world_size-2 gloo test of the problem-axis sharding and the final gather (CPU; the per-shard core is the
numpy oracle, injected through the `core` hook - the device engine needs a GPU)."""
import os
import socket
import warnings

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import pyjjasim_b200 as pj
from tests import cases


def _oracle_core(prob, th_mask, I_mask, w0, w1):
    from oracle import oracle
    c = prob.get_circuit()
    W = w1 - w0

    def cut(x, N):
        return np.broadcast_to(np.asarray(x), (N, prob.get_problem_count(), prob._Nt()))[:, w0:w1, :]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th, I = oracle.time_evolution_core(
            c.get_cycle_matrix(), c._Ic(), c._R(), c._C(), c._L(), prob._dt(), prob._Nt(), W,
            f=cut(prob.external_flux, c._Nf()), Is=cut(prob.current_sources, c._Nj()),
            Vs=cut(prob.voltage_sources, c._Nj()), T=cut(prob.temperature, c._Nj()),
            theta_m1=prob.config_at_minus_1[:, w0:w1], theta_m2=prob.config_at_minus_2[:, w0:w1],
            th_store_mask=th_mask, I_store_mask=I_mask)
    return th, I


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyjjasim_b200.distributed import compute_sharded, shard_for_rank
    kw, _ = cases.build("sq_frustrated", pj)
    prob = pj.TimeEvolutionProblem(**kw)
    res = compute_sharded(prob, core=_oracle_core)
    ret[rank] = (res.theta, res.current, res.voltage, shard_for_rank(prob.get_problem_count(), rank, world))
    dist.destroy_process_group()


def test_sharded_compute_equals_full_batch(golden_dir):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    g = np.load(os.path.join(golden_dir, "sq_frustrated.npz"))
    assert ret[0][3] == (0, 4) and ret[1][3] == (4, 6)
    for rank in (0, 1):
        th, I, V, _ = ret[rank]
        assert np.max(np.abs(th - g["theta"])) <= 1e-11
        assert np.max(np.abs(I - g["current"])) <= 1e-11
        assert np.max(np.abs(V - g["voltage"])) <= 1e-9
