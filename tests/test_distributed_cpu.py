"""world_size-2 gloo test of the N>1 host logic (sharding + the ensemble-average reduce). CPU only."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oqb_b200.sharding import ensemble_average, shard_bounds


def test_shard_bounds_cover_batch():
    for total in (0, 1, 7, 64, 65536, 8192):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, nobs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)  # same stream on every rank: the "global" per-member observables
    obs = rng.standard_normal((total, nobs)) + 1j * rng.standard_normal((total, nobs))
    lo, hi = shard_bounds(total, rank, world)
    local = torch.from_numpy(obs[lo:hi].sum(axis=0))
    avg = ensemble_average(local, hi - lo)
    ref = obs.mean(axis=0)
    ok = np.allclose(avg.numpy(), ref, rtol=1e-13, atol=1e-13)
    dist.barrier()
    dist.destroy_process_group()
    out[rank] = bool(ok)


def test_ensemble_average_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, 37, 5, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_ensemble_average_single_process():
    x = torch.tensor([1.0 + 2j, 3.0 - 1j], dtype=torch.complex128)
    assert torch.allclose(ensemble_average(x * 4, 4), x)
