"""world_size-2 gloo tests (CPU) of the reductions that replace jVMC.mpi_wrapper's MPI calls
(reference tests/mpi_wrapper_test.py:24-40 run under `mpirun -n 2`)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import vmc_jax_b200.mpi_wrapper as mpi
        from vmc_jax_b200.stats import SampledObs
        mpi._refresh()
        assert mpi.rank == rank and mpi.commSize == world
        # reference tests/mpi_wrapper_test.py: mean / variance of data split over ranks
        data = np.arange(720 * 4, dtype=np.float64).reshape(720, 4)
        myN, globN = mpi.distribute_sampling(720)
        mine = torch.as_tensor(data[rank * myN:(rank + 1) * myN]).reshape(1, -1, 4)
        p = torch.ones(mine.shape[:2], dtype=torch.float64) / globN
        assert torch.allclose(mpi.global_mean(mine, p), torch.as_tensor(data.mean(0)), atol=1e-10)
        assert torch.allclose(mpi.global_variance(mine, p), torch.as_tensor(data.var(0)), atol=1e-9)
        assert float(mpi.global_sum(torch.ones(1, 5, 1))) == 5 * world
        # complex covariance + gather with ragged sizes
        rng = np.random.default_rng(5)
        full = rng.normal(size=(11, 3)) + 1j * rng.normal(size=(11, 3))
        lo, hi = (0, 4) if rank == 0 else (4, 11)
        loc = torch.as_tensor(full[lo:hi]).reshape(1, -1, 3)
        g = mpi.gather(loc)
        assert np.allclose(g.numpy(), full)
        assert mpi.gather_offset(hi - lo) == lo
        w = torch.ones((1, hi - lo), dtype=torch.float64) / 11
        cov = mpi.global_covariance(loc, w)
        assert np.allclose(cov.numpy(), full.conj().T @ full / 11)
        # SampledObs across ranks == single-process numpy (reference tests/stats_test.py scaled by commSize)
        obs = SampledObs(loc, w)
        mean = full.mean(0)
        assert np.allclose(obs.mean().numpy(), mean)
        d = (full - mean) / np.sqrt(11)
        assert np.allclose(obs.covar().numpy(), d.conj().T @ d)
        assert np.allclose(obs.var().numpy(), (np.abs(d) ** 2).sum(0))
        assert np.allclose(obs.tangent_kernel().numpy(), d @ d.conj().T)
        sub = obs.subset(start=0, step=2)
        fs = np.concatenate([full[0:4][::2], full[4:11][::2]])
        ms = fs.mean(0)
        assert np.allclose(sub.mean().numpy(), ms)
        ds = (fs - ms) / np.sqrt(len(fs))
        assert np.allclose(sub.covar().numpy(), ds.conj().T @ ds)
        b = mpi.bcast_unknown_size(np.arange(7, dtype=np.float64) if rank == 0 else None)
        assert np.array_equal(b, np.arange(7.0))
        # sample distribution bookkeeping over two ranks
        spc, glob = mpi.distribute_sampling(1001, localDevices=1, numChainsPerDevice=10)
        assert glob == (51 + 50) * 10 and spc == (51 if rank == 0 else 50)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_two_rank_reductions_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", msg
