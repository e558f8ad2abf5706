"""CPU-only, world_size = 2 over gloo: the N > 1 host path (replica partition, gather of per-replica
observables to rank 0, max-over-ranks timing).  The data path itself has no collective."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, R, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import isingmodel_jl_b200 as pkg
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = pkg.sharding.replica_range(R, rank, world)
    local = np.stack([np.arange(lo, hi, dtype=np.float64), 10.0 * np.arange(lo, hi)], axis=1)
    allv = pkg.sharding.gather_replica_values(local, R)
    tmax = pkg.sharding.max_over_ranks(1.0 + rank)
    q.put((rank, None if allv is None else allv.tolist(), tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_and_timing():
    R, world, port = 7, 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, allv, tmax = q.get(timeout=120)
        res[rank] = (allv, tmax)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1][0] is None
    got = np.array(res[0][0])
    assert got.shape == (R, 2)
    assert np.array_equal(got[:, 0], np.arange(R)) and np.array_equal(got[:, 1], 10.0 * np.arange(R))
    assert res[0][1] == 2.0 and res[1][1] == 2.0
